// rbphd_analysis.cuh -- launchers of the kernels around the navigator's hot path:
// measurement generation of the simulated vehicle and the OSPA map metric of the post-analysis.
#pragma once
#include "rbphd_kernels.cuh"

namespace rbphd {

// SimulatedVehicle.Measure (SIMV:243-295) with the host's random numbers: per landmark one uniform (detection) and
// three N(0,1) draws (measurement noise, chol = lower Cholesky root of R), then nc clutter points from 3 uniforms each.
// z: (n + nc) x 3, assoc: n + nc (landmark index, INT_MIN for clutter), count: number of measurements written.
void launch_generate_measurements(cudaStream_t s, const DevCfg& cfg, const double* pose7, const double* landmarks, int n,
                                  const double* uniforms, const double* gauss, const double* chol9,
                                  const double* clutter_u, int nc, double* z, int* assoc, int* count);

// OSPA(a, b) (postanalysis/Plot.cs:531-581): optimal assignment of the thresholded distance matrix by a CTA-parallel
// Hungarian method on a dense nb x nb profit matrix in `work` (nb * nb + 8 * nb doubles).  out[0] = OSPA,
// out[1] = cardinality error.  na <= nb.
size_t ospa_workspace_doubles(int nb);
void launch_ospa(cudaStream_t s, const double* a, int na, const double* b, int nb, double c, double p, double* work,
                 double* out2);

}  // namespace rbphd
