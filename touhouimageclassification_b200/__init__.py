"""B200-native (sm_100a) implementation of TouhouIC's ViT training / inference hot path."""
__version__ = "0.1.0"
