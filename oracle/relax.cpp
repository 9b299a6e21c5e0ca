// placeholder until the RELAX restatement lands
#include "oracle.h"
int oracle_relax_dispatch(const char*, const void*, int, hlsl::Tex*, int, int, int) { return -1; }
