"""poreover_b200: B200-native backend for PoreOver's decoding hot path."""
__version__ = "0.1.0"
