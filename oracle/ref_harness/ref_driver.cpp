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

// ---- the reference's control panel driven headlessly (src/Kobayashi.cpp:383-629): one stand-in DX12App per object ----
static DX12App* gui_of(Kobayashi* k, bool create) {
    static std::vector<std::pair<Kobayashi*, DX12App*>> apps;
    for (auto& a : apps) if (a.first == k) return a.second;
    if (!create) return nullptr;
    DX12App* app = new DX12App();
    apps.push_back({k, app});
    return app;
}
static HWND__ g_panel;
// DX12App::initialize + WM_CREATE: iSetDXApp, iCreateObject, iWMCreate (creates the nine scrollbars)
void ref_gui_attach(void* h) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    DX12App* app = gui_of(k, true);
    app->setSimulation(k);
    k->iWMCreate(&g_panel, nullptr);
}
// WM_COMMAND with LOWORD(wParam) = com: 9 Reset, 10 Play/Pause, 11 Stop, 12 Next step (enum COM, src/Kobayashi.h:70-77)
void ref_gui_command(void* h, int com) { static_cast<Kobayashi*>(h)->iWMCommand(&g_panel, 0, (WPARAM)com, 0, nullptr); }
// WM_HSCROLL from slider `index` (0 tau .. 8 tEq): code = SB_* request, pos = thumb position for SB_THUMBTRACK
void ref_gui_hscroll(void* h, int index, int code, int pos) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    k->iWMHScroll(&g_panel, (WPARAM)((unsigned)code | ((unsigned)pos << 16)), (LPARAM)k->_crystalParameter[index].scrollbar, nullptr);
}
// one pass of the viewer's frame loop: DX12App::update (iUpdate when playing + all constant buffers) and draw
void ref_gui_frame(void* h) { DX12App* app = gui_of(static_cast<Kobayashi*>(h), false); if (app) { app->update(); app->draw(); } }
// colours the viewer currently holds, 3 per object
void ref_gui_colors(void* h, ref_real* rgb) {
    DX12App* app = gui_of(static_cast<Kobayashi*>(h), false);
    if (!app) return;
    for (size_t i = 0; i < app->_constantBuffer.size(); ++i) {
        rgb[3 * i + 0] = app->_constantBuffer[i].color.x; rgb[3 * i + 1] = app->_constantBuffer[i].color.y; rgb[3 * i + 2] = app->_constantBuffer[i].color.z;
    }
}
// out[0] playing, out[1] _simFrame, out[2..10] slider values (float members), out[11..19] slider integer positions
void ref_gui_state(void* h, double* out) {
    Kobayashi* k = static_cast<Kobayashi*>(h);
    out[0] = k->_updateFlag ? 1.0 : 0.0;
    out[1] = (double)k->_simFrame;
    for (int i = 0; i < 9; ++i) { out[2 + i] = (double)k->_crystalParameter[i].param_f.value; out[11 + i] = (double)k->_crystalParameter[i].param_i.value; }
}

}  // extern "C"
