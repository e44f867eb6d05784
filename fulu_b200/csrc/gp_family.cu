// One covariance family's kernels (fixed-theta LML, fit evaluation, prediction): compiled once per family with
// -DGP_FAM_F1=<product factor> -DGP_FAM_EX=<added term> (gp_block.cuh: GP_F_*), so the families build in parallel and the
// default family's code is untouched by the others.
#include "gp_device.cuh"

#ifndef GP_FAM_F1
#define GP_FAM_F1 0
#endif
#ifndef GP_FAM_EX
#define GP_FAM_EX 0
#endif

const GpFamilyOps* GP_FAMILY_SYMBOL(GP_FAM_F1, GP_FAM_EX)() { return GpFamilyLaunch<GpKern<GP_FAM_F1, GP_FAM_EX>>::ops(); }
