// oracle/detmath_export.cpp — TEST INFRASTRUCTURE: array entry points over detmath.h so that
// tests/test_detmath.py can bound it against libm on the CPU.
#include "../gfxexp_b200/csrc/detmath.h"
#include <cstddef>
extern "C" {
void orc_dm_sincos(const float* x, float* s, float* c, size_t n) { for (size_t i = 0; i < n; ++i) dm_sincos(x[i], &s[i], &c[i]); }
void orc_dm_acos(const float* x, float* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = dm_acos(x[i]); }
void orc_dm_atan2(const float* a, const float* b, float* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = dm_atan2(a[i], b[i]); }
void orc_dm_exp(const float* x, float* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = dm_exp(x[i]); }
void orc_dm_log(const float* x, float* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = dm_log(x[i]); }
void orc_dm_pow(const float* x, const float* e, float* y, size_t n) { for (size_t i = 0; i < n; ++i) y[i] = dm_pow(x[i], e[i]); }
}
