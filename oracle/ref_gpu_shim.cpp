// oracle/ref_gpu_shim.cpp -- TEST / BASELINE INFRASTRUCTURE ONLY.
// The reference's core.cu only DECLARES `int divup(int, int)` (core.h:74); the definition lives in the
// TA's utils.h:11-13, i.e. in main.cu's translation unit.  libref_gpu.so (core.cu alone, compiled
// unmodified for sm_100a) therefore needs it from somewhere: this is the ceil-division restated.
int divup(int a, int b) { return (a + b - 1) / b; }
