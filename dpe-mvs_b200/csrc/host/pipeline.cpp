// placeholder until the C++ host pipeline lands (next commit)
#include "../../../include/dpe_b200.h"
extern "C" __attribute__((visibility("default"))) int dpe_run_pipeline(const char*, int, int, int, int, int, int, int, int) { return DPE_ERR_STATE; }
