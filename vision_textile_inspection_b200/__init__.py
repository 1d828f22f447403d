"""B200-native per-frame inspection hot path."""
