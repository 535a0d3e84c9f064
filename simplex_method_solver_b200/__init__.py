"""B200-native (sm_100a) tableau pivot loop behind the reference's simplex.py surface."""
__version__ = "0.1.0"
