"""scan_b200: B200-native condgraph middle head (SCAN)."""
