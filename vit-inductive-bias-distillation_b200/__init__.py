"""B200-native BASD loss hot path."""
