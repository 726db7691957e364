"""B200-native ray-propagation engine for waveguide AR displays (hot path only)."""
__version__ = "0.1.0"
