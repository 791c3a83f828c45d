"""fvgn_b200: B200-native FVGN encode-process-decode step."""
