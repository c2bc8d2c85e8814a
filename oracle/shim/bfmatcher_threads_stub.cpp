// The GPU drop-in build has no CPU matcher; the harness's thread knobs are no-ops there.
extern "C" __attribute__((visibility("default"))) void plref_set_threads(int) {}
extern "C" __attribute__((visibility("default"))) int plref_get_threads() { return 0; }
