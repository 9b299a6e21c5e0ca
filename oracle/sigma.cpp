// placeholder until the SIGMA restatement lands
#include "oracle.h"
int oracle_sigma_dispatch(const char*, const void*, int, hlsl::Tex*, int, int, int) { return -1; }
