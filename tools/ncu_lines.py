"""Aggregate an ncu --page source --print-source cuda,sass --csv dump by CUDA source line: samples + instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
cur = None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 8 and r[0] not in ('', 'Line No') and r[0].isdigit():
        try: lines.append((cur, int(r[0]), r[1], int(r[6] or 0), int(r[7] or 0)))
        except ValueError: pass
tot = sum(l[3] for l in lines); toti = sum(l[4] for l in lines)
print("total samples", tot, "warp-instructions", toti)
byfile = collections.Counter(); byfilei = collections.Counter()
for f, n, s, smp, ins in lines: byfile[f] += smp; byfilei[f] += ins
for f, c in byfile.most_common(): print("  %-16s samples %6.2f%%  instr %6.2f%%" % (f, 100 * c / tot, 100 * byfilei[f] / toti))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for f, ln, s, smp, ins in sorted(lines, key=lambda l: -l[3])[:n]:
    print("%-14s %4d  smp %5.2f%%  ins %5.2f%%  %s" % (f, ln, 100 * smp / tot, 100 * ins / toti, s.strip()[:120]))
