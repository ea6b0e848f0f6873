"""furniture_b200: B200-native batched physics backend for the furniture-assembly env hot path."""
__version__ = "0.1.0"
