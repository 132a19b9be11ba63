// ref_driver.cpp — C entry points around the UNMODIFIED reference translation unit.
// The reference source is never copied: it is #included from where it lies (/root/reference/src, given
// with -I by oracle/Makefile) and compiled into oracle/_ref/ (git-ignored).  Test infrastructure only.
//
// -DKOB_REF_FP64 builds the "FP64-typed" oracle of SURVEY.md §8c: the same text with every `float`
// read as `double` (literals stay float-rounded, the dead-band stays FLT_EPSILON).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#ifdef KOB_REF_FP64
#define float double
#endif
#define private public
#include "Kobayashi.cpp"   // reference TU, in place
#undef private

#ifdef KOB_REF_FP64
typedef double ref_real;
#undef float
#else
typedef float ref_real;
#endif

extern "C" {

int ref_real_bytes(void) { return (int)sizeof(ref_real); }
void* ref_create(int nx, int ny, double dt) { return new Kobayashi(nx, ny, (ref_real)dt); }
void ref_destroy(void* h) { delete static_cast<Kobayashi*>(h); }
// which: 0 dx 1 dy 2 dt 3 tau 4 epsilonBar 5 mu 6 K 7 delta 8 anisotropy 9 alpha 10 gamma 11 tEq
void ref_set_param(void* h, int which, double v) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    ref_real* slot[] = {&k->_dx, &k->_dy, &k->_dt, &k->_tau, &k->_epsilonBar, &k->_mu, &k->_K, &k->_delta,
                        &k->_anisotropy, &k->_alpha, &k->_gamma, &k->_tEq};
    if (which >= 0 && which < 12) *slot[which] = (ref_real)v;
}
double ref_get_param(void* h, int which) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    ref_real* slot[] = {&k->_dx, &k->_dy, &k->_dt, &k->_tau, &k->_epsilonBar, &k->_mu, &k->_K, &k->_delta,
                        &k->_anisotropy, &k->_alpha, &k->_gamma, &k->_tEq};
    return (which >= 0 && which < 12) ? (double)*slot[which] : 0.0;
}
void ref_reset(void* h) { static_cast<Kobayashi*>(h)->_vectorInit(); }
void ref_step(void* h, int64_t n) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    for (int64_t s = 0; s < n; ++s) { k->_computeGradientLaplacian(); k->_evolution(); }
}
void ref_update(void* h) { static_cast<Kobayashi*>(h)->iUpdate(); }
void ref_get_fields(void* h, ref_real* phi, ref_real* t, ref_real* angl) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    const size_t b = sizeof(ref_real) * k->_phi.size();
    if (phi) std::memcpy(phi, k->_phi.data(), b);
    if (t) std::memcpy(t, k->_t.data(), b);
    if (angl) std::memcpy(angl, k->_angl.data(), b);
}
void ref_set_fields(void* h, const ref_real* phi, const ref_real* t, const ref_real* angl) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    const size_t b = sizeof(ref_real) * k->_phi.size();
    if (phi) std::memcpy(k->_phi.data(), phi, b);
    if (t) std::memcpy(k->_t.data(), t, b);
    if (angl) std::memcpy(k->_angl.data(), angl, b);
}
void ref_add_nucleus(void* h, int x, int y) { static_cast<Kobayashi*>(h)->_createNucleus(x, y); }
// Colour of object i through iUpdateConstantBuffer (src/Kobayashi.cpp:309-345); rgb = 3 values per object.
void ref_colors(void* h, ref_real* rgb) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    std::vector<ConstantBuffer> cb(k->_phi.size());
    for (int i = 0; i < (int)cb.size(); ++i) {
        k->iUpdateConstantBuffer(cb, i);
        rgb[3 * i + 0] = cb[i].color.x; rgb[3 * i + 1] = cb[i].color.y; rgb[3 * i + 2] = cb[i].color.z;
    }
}

}  // extern "C"
