"""B200-native reconstruction-distortion metrics (placeholder, filled in below)."""
