// TEST HARNESS ONLY: a stand-alone driver of the host-thread emulation (host_gp_block.cpp) meant to be built with
// -fsanitize=address,undefined (out-of-bounds indices into the "shared memory" block, which is allocated at exactly
// gp_smem_doubles(n); misaligned GpPair accesses; signed overflow) and with -fsanitize=thread (two "CUDA threads" touching
// the same shared-memory word without a barrier between them: what compute-sanitizer's racecheck reports, which is closed on
// the GPU pool).  Usage: sanitize_main <n> <nthr> <family> <use_kslab> <n_obs> [tiled]; prints lml and a checksum of the predictions.
// tiled = 1: the long-curve path (gp_tiled.cuh) - slab and stages allocated at exactly their sizes, the copy engine's memcpy against
// the warps' reads of the stages, two evaluations in a row so that every stage is refilled many times.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "host_gp_block.cpp"

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 100, nthr = argc > 2 ? atoi(argv[2]) : 128, family = argc > 3 ? atoi(argv[3]) : 0;
  const int kslab = argc > 4 ? atoi(argv[4]) : 0, n_obs = argc > 5 ? atoi(argv[5]) : 16, n_bands = 6;
  const int tiled = argc > 6 ? atoi(argv[6]) : 0;
  std::vector<double> t(n), flux(n), ferr(n), loglam(n_bands), xs(n), y(n), al(n), lam(n_bands), stats(6);
  std::vector<int32_t> band(n);
  unsigned long long r = 88172645463325252ull;
  auto u = [&] { r ^= r << 13; r ^= r >> 7; r ^= r << 17; return (double)(r >> 11) / 9007199254740992.0; };
  const double lams[6] = {3751.36, 4741.64, 6173.23, 7501.62, 8679.19, 9711.53};
  for (int b = 0; b < n_bands; ++b) loglam[b] = log10(lams[b]);
  for (int i = 0; i < n; ++i) {
    t[i] = 59000.0 + 200.0 * (i + u()) / n;
    band[i] = (int)(6 * u()) % 6;
    ferr[i] = 1.5 + 6.5 * u();
    const double x = (t[i] - 59060.0);
    flux[i] = 300.0 * exp(-x / 30.0) / (1.0 + exp(-x / 4.0)) * (1.0 + 0.15 * (lams[band[i]] / 6000.0 - 1.0)) + ferr[i] * (2.0 * u() - 1.0);
  }
  int st = host_gp_prepare(n, t.data(), flux.data(), ferr.data(), band.data(), loglam.data(), n_bands, n > 2 ? 1 : 0, 1e-10, nthr, xs.data(), y.data(),
                           al.data(), lam.data(), stats.data());
  const int fam = family & 0xff;   // (bits above are layout options of the harness: 0x400 = the scratch-free shared-memory layout)
  const int P = 4 + ((fam & 15) ? 1 : 0) + ((fam >> 4) == 4 ? 2 : ((fam >> 4) ? 1 : 0));
  std::vector<double> theta(P, 0.0), grad(P);
  theta[0] = 0.3; theta[P - 1] = -2.0;
  double lml = 0.0, lml2 = 0.0;
  int ok = tiled ? host_gp_lml_tiled(family, n, xs.data(), band.data(), y.data(), al.data(), lam.data(), theta.data(), nthr, 2, &lml, grad.data())
                 : host_gp_lml(family | (kslab ? 0x200 : 0), n, xs.data(), band.data(), y.data(), al.data(), lam.data(), theta.data(), nthr, &lml, grad.data());
  const int m = n_bands * n_obs;
  std::vector<double> mean(m), stdv(m);
  int ps = tiled ? host_gp_predict_tiled(family, n, xs.data(), band.data(), y.data(), al.data(), lam.data(), theta.data(), stats.data(), m, n_obs, t[0],
                                         t[n - 1], nullptr, nullptr, nthr, mean.data(), stdv.data(), &lml2)
                 : host_gp_predict(fam, n, xs.data(), band.data(), y.data(), al.data(), lam.data(), theta.data(), stats.data(), m, n_obs, t[0],
                                   t[n - 1], nullptr, nullptr, nthr, mean.data(), stdv.data(), &lml2);
  double sum = 0.0;
  for (int q = 0; q < m; ++q) sum += mean[q] + stdv[q];
  printf("n=%d nthr=%d family=%d kslab=%d prepare=%d ok=%d predict=%d lml=%.12g lml_predict=%.12g grad0=%.6g checksum=%.9g\n", n, nthr, family,
         kslab, st, ok, ps, lml, lml2, grad[0], sum);
  return (ok == 1 && std::isfinite(lml) && std::fabs(lml - lml2) <= 1e-9 * std::fabs(lml) && std::isfinite(sum)) ? 0 : 1;
}
