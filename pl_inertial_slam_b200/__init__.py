"""B200-native descriptor matching front-end of PL-inertial-slam."""
