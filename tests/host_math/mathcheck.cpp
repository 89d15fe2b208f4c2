// Host build of the fast-mode elementary functions (tfg_math.cuh compiles for the host as plain C++), exported
// as array functions so that tests/test_host_math.py can measure their error against long-double libm.
#include "../../topoflow_glacier_b200/csrc/tfg_math.cuh"

using namespace tfg::fm;

#define ARRAY_FN(name, expr)                                                   \
  extern "C" void name(const double* x, const double* y, double* out, long n) { \
    for (long i = 0; i < n; ++i) { const double a = x[i], b = y ? y[i] : 0.0; (void)b; out[i] = (expr); } \
  }
ARRAY_FN(mc_exp_core, exp_core(a))
ARRAY_FN(mc_log_core, log_core(a))
ARRAY_FN(mc_exp_tab, exp_tab(a))
ARRAY_FN(mc_log_tab, log_tab(a))
ARRAY_FN(mc_rcp, rcp(a))
ARRAY_FN(mc_rcp3, rcp3(a))
ARRAY_FN(mc_div, div(a, b))
ARRAY_FN(mc_div_fast, div_fast(a, b))
ARRAY_FN(mc_sqrt_pos, sqrt_pos(a))
ARRAY_FN(mc_asin01, asin01(a))
ARRAY_FN(mc_atan_core, atan_core(a))
ARRAY_FN(mc_stull, stull_wet_bulb(a, b))
ARRAY_FN(mc_atan_diff, atan_diff(a, b))
ARRAY_FN(mc_root7, root7(a, (float)b))
ARRAY_FN(mc_stull_tab, stull_wet_bulb_tab(a, b))
ARRAY_FN(mc_inv_air_mass, inv_air_mass(a))
// the N-at-once variants must return exactly what the scalar routines return
extern "C" void mc_exp_tab_2(const double* x, const double* y, double* out, long n) {
  for (long i = 0; i + 1 < n; i += 2) { const double a[2] = {x[i], x[i + 1]}; double r[2]; exp_tab_n<2>(a, r); out[i] = r[0]; out[i + 1] = r[1]; }
}
extern "C" void mc_log_tab_2(const double* x, const double* y, double* out, long n) {
  for (long i = 0; i + 1 < n; i += 2) { const double a[2] = {x[i], x[i + 1]}; double r[2]; log_tab_n<2>(a, r); out[i] = r[0]; out[i + 1] = r[1]; }
}
extern "C" void mc_rcp3_3(const double* x, const double* y, double* out, long n) {
  for (long i = 0; i + 2 < n; i += 3) { const double a[3] = {x[i], x[i + 1], x[i + 2]}; double r[3]; rcp3_n<3>(a, r); out[i] = r[0]; out[i + 1] = r[1]; out[i + 2] = r[2]; }
}
// the fused exp pair + 7th root must return exactly what the separate routines return
extern "C" void mc_exp2_root7(const double* x, const double* y, double* out, long n) {
  for (long i = 0; i + 2 < n; i += 3) {
    const double a[2] = {x[i], x[i + 1]}; double r[2], r7;
    exp_tab2_root7(a, r, y[i + 2], r7);
    out[i] = r[0]; out[i + 1] = r[1]; out[i + 2] = r7;
  }
}
