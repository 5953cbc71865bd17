// Temporary: entry points not yet implemented return MCOSB_ERR_STATE. Removed as each stage lands.
#include "common.cuh"
#define NOT_YET(ctx, what) return mcosb_fail(ctx, MCOSB_ERR_STATE, what " is not implemented yet")
extern "C" int mcosb_nco(mcosb_ctx* ctx, int, const double*, const double*, int, int, const int*, double*, int*, int*) { NOT_YET(ctx, "mcosb_nco"); }
int mcosb_dev_nco(mcosb_ctx* ctx, int, const double*, const double*, int, int, const int*, double*, int*, int*) { NOT_YET(ctx, "mcosb_nco"); }
