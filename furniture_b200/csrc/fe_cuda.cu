// fe_cuda.cu -- sm_100a kernels and the CUDA platform layer of the C-ABI (include/furniture_b200.h).
// One block = one warp = one environment; the warp's working set lives in dynamic shared memory for the whole call
// (all nsub mj_steps of an env step run without touching HBM except for the model tables, which stay in L1/L2).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define PLAT_IS_CUDA 1
struct fe_handle;
static void* plat_alloc(size_t bytes) { void* p = nullptr; return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr; }
static void plat_free(void* p) { cudaFree(p); }
static void plat_memset0(void* p, size_t n) { cudaMemset(p, 0, n); }
static void plat_upload(void* d, const void* h, size_t n) { cudaMemcpy(d, h, n, cudaMemcpyHostToDevice); }
static void plat_download(void* h, const void* d, size_t n) { cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost); }
static int plat_init(fe_handle* h);
static void plat_fini(fe_handle* h);
static void plat_sync(fe_handle* h);
static int plat_run_sim(fe_handle* h, int nsub, int mode, void* stream);
static int plat_run_reset(fe_handle* h, const uint8_t* mask, void* stream);
static int plat_run_step(fe_handle* h, const float* actions, float* reward, uint8_t* done, int32_t* info, void* stream);
static int plat_step_host(fe_handle* h, const float* actions, float* obs, float* reward, uint8_t* done, int32_t* info);
static void plat_copy_d2d(fe_handle* h, void* dst, const void* src, size_t n, void* stream);
static int plat_is_aligned(fe_handle* h, int n, const double* p1, const double* m1, const double* p2, const double* m2, const double* angles,
                           const int32_t* nangles, const double* thr, uint8_t* aligned, double* tq);

#include "fe_api.inl"

// ---------------------------------------------------------------- kernels
extern __shared__ float fe_smem[];

#define FE_MAX_WPB 14
__global__ void __launch_bounds__(32 * FE_MAX_WPB) fe_sim_kernel(FeState s, const fe_model* __restrict__ m, FeOpt opt, int nsub, int mode, FeDebug dbg, int slice_words) {
  const int wib = threadIdx.x >> 5, slot = blockIdx.x * (blockDim.x >> 5) + wib;
  if (slot >= s.N) return;
  const int env = slot;
  fe_run_env(s, m, opt, env, nsub, mode, fe_smem + (size_t)wib * slice_words, dbg);
}

__global__ void __launch_bounds__(32 * FE_MAX_WPB) fe_env_step_kernel(FeState st, FeEnvState es, const fe_model* __restrict__ m, const fe_scene* __restrict__ sc,
                                                         fe_config cfg, FeOpt opt, const float* __restrict__ actions, float* reward, uint8_t* done,
                                                         int32_t* info, int slice_words, const int* __restrict__ slots) {
  const int wib = threadIdx.x >> 5, slot = blockIdx.x * (blockDim.x >> 5) + wib;
  const int env = slots[slot];
  if (env < 0) return; // unused slot (blocks of heavy envs are deliberately left partly empty)
#ifdef FE_EXP_MODEL_SMEM
  if (blockDim.x == 32) { // experiment: one warp per block, model tables staged in shared memory behind the slice
    float* ms = fe_smem + (slice_words + FE_ENV_EXTRA_WORDS);
    const float* src = (const float*)m;
    for (int i = threadIdx.x; i < (int)(sizeof(fe_model) / 4); i += 32) ms[i] = src[i];
    __syncwarp();
    m = (const fe_model*)ms;
  }
#endif
  FeEnv e;
  fe_env_bind(&e, fe_smem + (size_t)wib * (slice_words + FE_ENV_EXTRA_WORDS), m, sc, &cfg, opt, st, es, env, slice_words);
  fe_load(e.w, st, env);
  fe_env_load_groups(&e);
  fe_env_step_one(&e, actions, reward, done, info);
  fe_env_store_groups(&e);
  fe_store(e.w, st, env);
}

__global__ void __launch_bounds__(32 * FE_MAX_WPB) fe_env_reset_kernel(FeState st, FeEnvState es, const fe_model* __restrict__ m, const fe_scene* __restrict__ sc,
                                                          fe_config cfg, FeOpt opt, const uint8_t* __restrict__ mask, int slice_words) {
  const int wib = threadIdx.x >> 5, env = blockIdx.x * (blockDim.x >> 5) + wib;
  if (env >= st.N) return;
  if (mask && !mask[env]) return;
  FeEnv e;
  fe_env_bind(&e, fe_smem + (size_t)wib * (slice_words + FE_ENV_EXTRA_WORDS), m, sc, &cfg, opt, st, es, env, slice_words);
  fe_load(e.w, st, env);
  fe_env_load_groups(&e);
  fe_env_reset_one(&e);
  fe_env_store_groups(&e);
  fe_store(e.w, st, env);
}

// Packs the envs into blocks for the next step from the work their last step took (cycles in the five phases, barrier
// waits excluded).  Counting sort on a log-scale key (16 buckets per octave), heaviest first, into order[].  The warps of a
// block run in lockstep, so like goes with like; and the few envs far heavier than the median (robot coupled to a part:
// the big Newton solve) bound the whole step by their own latency, which is lowest when few warps share the SM: they get
// blocks with only `heavy_k` of the warp slots used, launched first, while the light envs fill the other SMs.
#define FE_ORDER_BUCKETS 256
#define FE_EXTRA_BLOCKS 148
__global__ void __launch_bounds__(1024) fe_order_kernel(int N, const int* __restrict__ stats, int* __restrict__ order, int* __restrict__ slots, int nslots,
                                                        int wpb, int heavy_k, int heavy_shift, float* __restrict__ pred, float decay) {
  __shared__ int hist[FE_ORDER_BUCKETS], start[FE_ORDER_BUCKETS], nheavy;
  const int tid = threadIdx.x;
  if (tid < FE_ORDER_BUCKETS) hist[tid] = 0;
  // predicted work of the next step: the last step's, but an env that was heavy a few steps ago is still suspect
  for (int e = tid; e < N; e += 1024) {
    const int* st = stats + (size_t)e * FE_NSTAT;
    const float work = (float)st[4] + (float)st[5] + (float)st[6] + (float)st[7] + (float)st[8]; // cycles / 16
    pred[e] = fmaxf(work, decay * pred[e]);
  }
  __syncthreads();
  auto bucket_of = [&](int e) {
    const float work = pred[e];
    int b = (int)(16.f * log2f(fmaxf(work, 1.f) * (1.f / 1024.f)));                               // bucket 0 below 16k cycles
    b = b < 0 ? 0 : (b > FE_ORDER_BUCKETS - 1 ? FE_ORDER_BUCKETS - 1 : b);
    return FE_ORDER_BUCKETS - 1 - b; // heaviest first
  };
  for (int e = tid; e < N; e += 1024) atomicAdd(&hist[bucket_of(e)], 1);
  __syncthreads();
  if (tid == 0) {
    int acc = 0, med = -1;
    for (int b = 0; b < FE_ORDER_BUCKETS; ++b) { start[b] = acc; acc += hist[b]; if (med < 0 && 2 * acc >= N) med = b; }
    // heavy: at least heavy_shift buckets (sixteenths of an octave) above the median bucket
    const int hb = med - heavy_shift; // last heavy bucket (buckets are in heaviest-first order)
    const int H = (heavy_k > 0 && heavy_k < wpb && hb >= 0) ? start[hb] + hist[hb] : 0;
    const int cap = heavy_k * FE_EXTRA_BLOCKS;
    nheavy = H > cap ? cap : H;
  }
  __syncthreads();
  for (int e = tid; e < N; e += 1024) order[atomicAdd(&start[bucket_of(e)], 1)] = e;
  for (int i = tid; i < nslots; i += 1024) slots[i] = -1;
  __syncthreads();
  const int H = nheavy, HB = heavy_k > 0 ? (H + heavy_k - 1) / heavy_k : 0;
  for (int r = tid; r < N; r += 1024) {
    const int slot = r < H ? (r / heavy_k) * wpb + r % heavy_k : HB * wpb + (r - H);
    slots[slot] = order[r];
  }
}

__global__ void fe_is_aligned_kernel(int n, const double* p1, const double* m1, const double* p2, const double* m2, const double* cs, const double* sn,
                                     const int32_t* nang, const double* thr, uint8_t* aligned, double* tq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double q[4] = {0, 0, 0, 0};
  bool set = false;
  const bool ok = fe_is_aligned_d(p1 + 3 * i, m1 + 9 * i, p2 + 3 * i, m2 + 9 * i, nang[i], cs + 4 * i, sn + 4 * i, thr + 4 * i, q, &set);
  aligned[i] = ok ? 1 : 0;
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  for (int k = 0; k < 4; ++k) tq[4 * i + k] = set ? q[k] : nanv;
}

// ---------------------------------------------------------------- platform layer
struct CudaPlat {
  size_t smem_sim = 0, smem_env = 0;
  int wpb = 1;
  int reorder = 1, heavy_k = 7, heavy_shift = 18; // heavy: 2^(18/16) = 2.2x the median work; 7 of the 14 warp slots used
  int* slots = nullptr;  // block slot -> env (or -1)
  float* pred = nullptr; // per env: predicted work of the next step
  float decay = 0.85f;
  int nblocks = 0;
  float* pin_act = nullptr;
  unsigned char* pin_out = nullptr;
  size_t out_bytes = 0;
  cudaStream_t stream = nullptr;
  FeLayout* lay_pin = nullptr;
};
#define CUDA_OK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(h, -10, std::string(#call) + ": " + cudaGetErrorString(_e)); } while (0)

static int plat_init(fe_handle* h) {
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) { h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return -10; }
  CudaPlat* p = new CudaPlat();
  h->plat = p;
  e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking); // private stream of fe_env_step_host
  if (e != cudaSuccess) { h->err = std::string("cudaStreamCreateWithFlags: ") + cudaGetErrorString(e); delete p; h->plat = nullptr; return -10; }
  return 0;
}
// The slice layout table is one __constant__ object per process: (re)upload it when the handle about to launch uses a
// different layout than the one resident (several handles with different models alive at once: the tests, a mixed-furniture
// batch).  If every launch since the last upload went to the stream of this launch, the switch is one stream-ordered
// cudaMemcpyToSymbolAsync from the handle's pinned copy (kernels of the old layout are ahead of it in the same stream);
// otherwise the device is drained first.
static FeLayout g_resident_lay;
static bool g_resident_valid = false;
static std::vector<cudaStream_t> g_resident_streams; // streams that received launches under the resident layout
static int plat_use_layout(fe_handle* h, cudaStream_t stream) {
  CudaPlat* p = (CudaPlat*)h->plat;
  const bool same = g_resident_valid && memcmp(&g_resident_lay, &h->lay, sizeof(FeLayout)) == 0;
  if (!same) {
    const bool ordered = g_resident_valid && p->lay_pin && g_resident_streams.size() == 1 && g_resident_streams[0] == stream;
    if (ordered) {
      CUDA_OK(cudaMemcpyToSymbolAsync(fe_c_lay, p->lay_pin, sizeof(FeLayout), 0, cudaMemcpyHostToDevice, stream));
    } else {
      CUDA_OK(cudaDeviceSynchronize()); // kernels of the previous layout's handle must have drained
      CUDA_OK(cudaMemcpyToSymbol(fe_c_lay, &h->lay, sizeof(FeLayout)));
    }
    g_resident_lay = h->lay;
    g_resident_valid = true;
    g_resident_streams.clear();
  }
  bool seen = false;
  for (cudaStream_t s : g_resident_streams) seen |= s == stream;
  if (!seen) g_resident_streams.push_back(stream);
  return 0;
}
static int plat_prepare(fe_handle* h, cudaStream_t stream) {
  CudaPlat* p = (CudaPlat*)h->plat;
  if (!p->lay_pin) { // pinned copy of the layout: source of the stream-ordered switch
    CUDA_OK(cudaMallocHost((void**)&p->lay_pin, sizeof(FeLayout)));
    *p->lay_pin = h->lay;
  }
  if (int rc = plat_use_layout(h, stream)) return rc;
  if (p->smem_sim) return 0;
  // warps (= envs) per block: as many as fit in 227 KB of shared memory, at most FE_MAX_WPB; FE_WPB overrides
  const size_t per_env = (size_t)(h->slice_words + FE_ENV_EXTRA_WORDS) * 4;
  int wpb = (int)((227 * 1024 - 1024) / per_env);
  if (wpb > FE_MAX_WPB) wpb = FE_MAX_WPB;
  if (const char* e = getenv("FE_WPB")) { int v = atoi(e); if (v >= 1 && v <= wpb) wpb = v; }
  if (wpb < 1) return fail(h, -11, "model does not fit in shared memory");
  p->wpb = wpb;
  if (const char* e = getenv("FE_REORDER")) p->reorder = atoi(e);
  if (const char* e = getenv("FE_HEAVY_K")) p->heavy_k = atoi(e);
  if (const char* e = getenv("FE_HEAVY_SHIFT")) p->heavy_shift = atoi(e);
  p->nblocks = (h->N + wpb - 1) / wpb + FE_EXTRA_BLOCKS;
  {
    std::vector<int> init((size_t)p->nblocks * wpb, -1);
    for (int i = 0; i < h->N; ++i) init[i] = i;
    if (const char* e = getenv("FE_PRED_DECAY")) p->decay = 0.01f * (float)atoi(e);
    CUDA_OK(cudaMalloc((void**)&p->pred, sizeof(float) * (size_t)h->N));
    CUDA_OK(cudaMemset(p->pred, 0, sizeof(float) * (size_t)h->N));
    CUDA_OK(cudaMalloc((void**)&p->slots, sizeof(int) * init.size()));
    CUDA_OK(cudaMemcpy(p->slots, init.data(), sizeof(int) * init.size(), cudaMemcpyHostToDevice));
  }
  p->smem_sim = (size_t)h->slice_words * 4 * wpb;
  p->smem_env = per_env * wpb;
  // the opt-in limit is an attribute of the kernel, not of the handle: always the hardware maximum (227 KB), so that handles
  // with different slice sizes can be alive together (a smaller value set by a later handle would fail the earlier one's launches)
  CUDA_OK(cudaFuncSetAttribute(fe_sim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CUDA_OK(cudaFuncSetAttribute(fe_env_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CUDA_OK(cudaFuncSetAttribute(fe_env_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const size_t N = h->N;
  p->out_bytes = N * (sizeof(float) * h->hs.obs_dim + sizeof(float) + sizeof(int32_t) * FE_INFO_DIM + 1);
  CUDA_OK(cudaMallocHost((void**)&p->pin_act, sizeof(float) * N * (h->hs.act_dim > 0 ? h->hs.act_dim : 1)));
  CUDA_OK(cudaMallocHost((void**)&p->pin_out, p->out_bytes + 64));
  return 0;
}
static void plat_fini(fe_handle* h) {
  CudaPlat* p = (CudaPlat*)h->plat;
  if (!p) return;
  cudaDeviceSynchronize();
  if (p->lay_pin) cudaFreeHost(p->lay_pin);
  if (p->pin_act) cudaFreeHost(p->pin_act);
  if (p->pin_out) cudaFreeHost(p->pin_out);
  if (p->slots) cudaFree(p->slots);
  if (p->pred) cudaFree(p->pred);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  h->plat = nullptr;
}
static void plat_sync(fe_handle* h) { cudaDeviceSynchronize(); }
static void plat_copy_d2d(fe_handle* h, void* dst, const void* src, size_t n, void* stream) {
  if (dst != src) cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}
static int plat_run_sim(fe_handle* h, int nsub, int mode, void* stream) {
  int rc = plat_prepare(h, (cudaStream_t)stream);
  if (rc) return rc;
  CudaPlat* p = (CudaPlat*)h->plat;
  fe_sim_kernel<<<(h->N + p->wpb - 1) / p->wpb, 32 * p->wpb, p->smem_sim, (cudaStream_t)stream>>>(h->st, h->dm, h->opt, nsub, mode, h->dbg, h->slice_words);
  CUDA_OK(cudaGetLastError());
  return 0;
}
static int plat_run_reset(fe_handle* h, const uint8_t* mask, void* stream) {
  int rc = plat_prepare(h, (cudaStream_t)stream);
  if (rc) return rc;
  CudaPlat* p = (CudaPlat*)h->plat;
  fe_env_reset_kernel<<<(h->N + p->wpb - 1) / p->wpb, 32 * p->wpb, p->smem_env, (cudaStream_t)stream>>>(h->st, h->es, h->dm, h->ds, h->cfg, h->opt, mask, h->slice_words);
  CUDA_OK(cudaGetLastError());
  return 0;
}
static int plat_run_step(fe_handle* h, const float* actions, float* reward, uint8_t* done, int32_t* info, void* stream) {
  int rc = plat_prepare(h, (cudaStream_t)stream);
  if (rc) return rc;
  CudaPlat* p = (CudaPlat*)h->plat;
#ifdef FE_EXP_MODEL_SMEM
  fe_env_step_kernel<<<p->nblocks, 32 * p->wpb, p->wpb == 1 ? p->smem_env + sizeof(fe_model) : p->smem_env, (cudaStream_t)stream>>>(h->st, h->es, h->dm, h->ds, h->cfg, h->opt, actions, reward, done, info, h->slice_words, p->slots);
#else
  fe_env_step_kernel<<<p->nblocks, 32 * p->wpb, p->smem_env, (cudaStream_t)stream>>>(h->st, h->es, h->dm, h->ds, h->cfg, h->opt, actions, reward, done, info, h->slice_words, p->slots);
#endif
  CUDA_OK(cudaGetLastError());
  if (p->reorder) fe_order_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(h->N, h->st.stats, h->st.order, p->slots, p->nblocks * p->wpb, p->wpb, p->heavy_k, p->heavy_shift, p->pred, p->decay);
  return 0;
}
static int plat_step_host(fe_handle* h, const float* actions, float* obs, float* reward, uint8_t* done, int32_t* info) {
  CudaPlat* p = (CudaPlat*)h->plat;
  int rc = plat_prepare(h, p->stream);
  if (rc) return rc;
  const size_t N = h->N, ab = sizeof(float) * N * h->hs.act_dim, ob = sizeof(float) * N * h->hs.obs_dim, rb = sizeof(float) * N, ib = sizeof(int32_t) * N * FE_INFO_DIM;
  memcpy(p->pin_act, actions, ab);
  // the private stream does not order against work the caller issued on other streams (fe_sim_forward, fe_set_field ...)
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpyAsync(h->dev_act, p->pin_act, ab, cudaMemcpyHostToDevice, p->stream));
  fe_env_step_kernel<<<p->nblocks, 32 * p->wpb, p->smem_env, p->stream>>>(h->st, h->es, h->dm, h->ds, h->cfg, h->opt, (const float*)h->dev_act, (float*)h->dev_rew,
                                                            (uint8_t*)h->dev_done, (int32_t*)h->dev_info, h->slice_words, p->slots);
  CUDA_OK(cudaGetLastError());
  if (p->reorder) fe_order_kernel<<<1, 1024, 0, p->stream>>>(h->N, h->st.stats, h->st.order, p->slots, p->nblocks * p->wpb, p->wpb, p->heavy_k, p->heavy_shift, p->pred, p->decay);
  unsigned char* o = p->pin_out;
  CUDA_OK(cudaMemcpyAsync(o, h->es.obs, ob, cudaMemcpyDeviceToHost, p->stream));
  CUDA_OK(cudaMemcpyAsync(o + ob, h->dev_rew, rb, cudaMemcpyDeviceToHost, p->stream));
  CUDA_OK(cudaMemcpyAsync(o + ob + rb, h->dev_info, ib, cudaMemcpyDeviceToHost, p->stream));
  CUDA_OK(cudaMemcpyAsync(o + ob + rb + ib, h->dev_done, N, cudaMemcpyDeviceToHost, p->stream));
  CUDA_OK(cudaStreamSynchronize(p->stream));
  if (obs) memcpy(obs, o, ob);
  if (reward) memcpy(reward, o + ob, rb);
  if (info) memcpy(info, o + ob + rb, ib);
  if (done) memcpy(done, o + ob + rb + ib, N);
  return 0;
}
static int plat_is_aligned(fe_handle* h, int n, const double* p1, const double* m1, const double* p2, const double* m2, const double* angles,
                           const int32_t* nangles, const double* thr, uint8_t* aligned, double* tq) {
  // cos/sin of the allowed angles are evaluated on the host in float64 (the same libm the reference's numpy uses)
  std::vector<double> cs(4 * (size_t)n), sn(4 * (size_t)n);
  for (size_t i = 0; i < 4 * (size_t)n; ++i) { double a = angles[i] / 180 * 3.141592653589793; cs[i] = cos(a); sn[i] = sin(a); }
  double *d_p1, *d_m1, *d_p2, *d_m2, *d_cs, *d_sn, *d_thr, *d_tq;
  int32_t* d_na;
  uint8_t* d_al;
  const size_t N = n;
  CUDA_OK(cudaMalloc(&d_p1, 24 * N)); CUDA_OK(cudaMalloc(&d_m1, 72 * N)); CUDA_OK(cudaMalloc(&d_p2, 24 * N)); CUDA_OK(cudaMalloc(&d_m2, 72 * N));
  CUDA_OK(cudaMalloc(&d_cs, 32 * N)); CUDA_OK(cudaMalloc(&d_sn, 32 * N)); CUDA_OK(cudaMalloc(&d_thr, 32 * N)); CUDA_OK(cudaMalloc(&d_tq, 32 * N));
  CUDA_OK(cudaMalloc(&d_na, 4 * N)); CUDA_OK(cudaMalloc(&d_al, N));
  cudaMemcpy(d_p1, p1, 24 * N, cudaMemcpyHostToDevice); cudaMemcpy(d_m1, m1, 72 * N, cudaMemcpyHostToDevice);
  cudaMemcpy(d_p2, p2, 24 * N, cudaMemcpyHostToDevice); cudaMemcpy(d_m2, m2, 72 * N, cudaMemcpyHostToDevice);
  cudaMemcpy(d_cs, cs.data(), 32 * N, cudaMemcpyHostToDevice); cudaMemcpy(d_sn, sn.data(), 32 * N, cudaMemcpyHostToDevice);
  cudaMemcpy(d_thr, thr, 32 * N, cudaMemcpyHostToDevice); cudaMemcpy(d_na, nangles, 4 * N, cudaMemcpyHostToDevice);
  fe_is_aligned_kernel<<<(n + 127) / 128, 128>>>(n, d_p1, d_m1, d_p2, d_m2, d_cs, d_sn, d_na, d_thr, d_al, d_tq);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpy(aligned, d_al, N, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(tq, d_tq, 32 * N, cudaMemcpyDeviceToHost));
  cudaFree(d_p1); cudaFree(d_m1); cudaFree(d_p2); cudaFree(d_m2); cudaFree(d_cs); cudaFree(d_sn); cudaFree(d_thr); cudaFree(d_tq); cudaFree(d_na); cudaFree(d_al);
  return 0;
}
