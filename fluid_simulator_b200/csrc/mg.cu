// mg.cu -- multigrid preconditioner for the pressure PCG (placeholder: disabled; diagonal scaling is used).
#include "fsim_internal.h"

bool mg_enabled(const fsim* h) { (void)h; return false; }
int mg_build(fsim* h) { (void)h; return FSIM_OK; }
int mg_apply(fsim* h) { (void)h; return FSIM_OK; }
void mg_free(fsim* h) { (void)h; }
