// KobayashiSimulation.hpp — DXViewer plugin adapter: implements the viewer's `ISimulation` interface
// (ext/DXViewer/DXViewer-3.1.0/include/ISimulation.h:7-85, 20 pure virtuals) on top of the GPU-backed host class
// `Kobayashi` (Kobayashi.hpp -> kobayashi_c.h -> libkobayashi_cuda.so), so that the reference application can do
//
//     auto* sim = new KobayashiSimulation(250, 250, 0.0001f);      // src/main.cpp:14-18 constructs `Kobayashi` here
//     DX12App* dxapp = new DX12App();  dxapp->setSimulation(sim);
//
// and keep its viewer unchanged.  What the viewer sees is the same as with the reference class:
//   * one unit quad per cell (iGetVertices/iGetIndices/iCreateObject: src/Kobayashi.cpp:253-307),
//   * per frame `iUpdate()` = 10 fused sub-steps on the device (:227-239), then one colour per object from phi
//     through the 4-colour ramp (:309-345) — computed by the device kernel `kob_render_rgba`, read back once per
//     frame, and indexed with the reference's TRANSPOSED object -> cell mapping (object i shows cell
//     (x, y) = (i / n, i % n), :312-315),
//   * the control panel (iWM* hooks, :383-629) with the reference's semantics: the nine sliders of :17-58 (float value +
//     integer thumb position, same ranges / strides, the float accumulated by += stride exactly like the reference does),
//     a slider move writes the parameter and resets the simulation (:589-616), Reset restores the defaults of :76-96 and
//     resets (:512-528), Play toggles iUpdate on/off (:531-538), Stop resets (:540-543), Next step runs one iUpdate and
//     refreshes the viewer (:544-549); iResetSimulationState re-seeds, lets the viewer refresh (which, while playing, already
//     runs one iUpdate) and then zeroes the counters (:241-249).  Only the widgets themselves are not created.
//
// Include the viewer's "Win32App.h" (or any header that declares ISimulation, Vertex, ConstantBuffer, DX12App and the
// DirectX storage types) BEFORE this file.  Header only; no Direct3D call is made here except the two the interface
// forces in iDraw.
#ifndef KOBAYASHI_SIMULATION_HPP
#define KOBAYASHI_SIMULATION_HPP

#include <cmath>
#include <cstdint>
#include <vector>

#include "Kobayashi.hpp"

class KobayashiSimulation : public ISimulation {
public:
    // command ids of the control panel, enum COM of src/Kobayashi.h:70-77
    enum COM { TAU, EPLSILONBAR, MU, K, DELTA, ANISOTROPY, ALPHA, GAMMA, TEQ, RESET, PLAY, STOP, NEXTSTEP, TIME_TEXT, FRAME_TEXT };

    KobayashiSimulation(int x, int y, float timeStep, int device = 0, int kernel = KOB_KERNEL_FAST)
        : sim_(x, y, timeStep, KOB_F32, kernel, device), nx_(x), ny_(y) {
        quad_ = {Vertex{DirectX::XMFLOAT3(-0.5f, -0.5f, 0.0f)}, Vertex{DirectX::XMFLOAT3(-0.5f, +0.5f, 0.0f)},
                 Vertex{DirectX::XMFLOAT3(+0.5f, +0.5f, 0.0f)}, Vertex{DirectX::XMFLOAT3(+0.5f, -0.5f, 0.0f)}};
        tris_ = {0u, 1u, 2u, 0u, 2u, 3u};
        //            value     min      max      stride   | thumb: value min max stride | value = thumb * ratio     (:17-58)
        sliders_ = {{0.0003f, 0.0001f, 0.0009f, 0.0001f, 3, 1, 9, 1, 0.0001f},      // tau
                    {0.010f, 0.006f, 0.015f, 0.001f, 10, 6, 15, 1, 0.001f},         // epsilonBar
                    {1.0f, 0.5f, 1.4f, 0.1f, 10, 5, 14, 1, 0.1f},                   // mu (a slider, but unused by the model)
                    {1.6f, 1.0f, 1.9f, 0.1f, 16, 10, 19, 1, 0.1f},                  // K
                    {0.05f, 0.01f, 0.09f, 0.01f, 5, 1, 9, 1, 0.01f},                // delta
                    {6.0f, 2.0f, 8.0f, 1.0f, 6, 2, 8, 1, 1.0f},                     // anisotropy
                    {0.9f, 0.7f, 1.2f, 0.1f, 9, 7, 12, 1, 0.1f},                    // alpha
                    {10.0f, 10.0f, 20.0f, 1.0f, 10, 10, 20, 1, 1.0f},               // gamma
                    {1.0f, 0.5f, 1.5f, 0.1f, 10, 5, 15, 1, 0.1f}};                  // tEq
        parameterInit();
        refreshColours();
    }

    Kobayashi& sim() { return sim_; }

    // ---- simulation ----
    void iUpdate() override {
        sim_.nextStep();                      // 10 sub-steps + _simFrame/_simTime (:227-239); the caller checks iIsUpdated()
        refreshColours();
    }
    // _vectorInit, then the viewer refreshes (DX12App::update runs one iUpdate while playing), then the counters are zeroed
    void iResetSimulationState(std::vector<ConstantBuffer>& constantBuffer) override {
        sim_.iResetSimulationState();
        refreshColours();
        if (dxapp_) { dxapp_->update(); dxapp_->draw(); }
        else for (size_t i = 0; i < constantBuffer.size(); ++i) iUpdateConstantBuffer(constantBuffer, static_cast<int>(i));
        sim_.setSimCounters(0, 0.0);
    }
    bool iIsUpdated() override { return updateFlag_; }
    void setUpdated(bool play) { updateFlag_ = play; }
    void nextStep() { iUpdate(); }

    // ---- mesh: one unit quad, instanced once per cell ----
    std::vector<Vertex>& iGetVertices() override { return quad_; }
    std::vector<unsigned int>& iGetIndices() override { return tris_; }
    UINT iGetVertexBufferSize() override { return static_cast<UINT>(quad_.size()); }
    UINT iGetIndexBufferSize() override { return static_cast<UINT>(tris_.size()); }
    DirectX::XMINT3 iGetObjectCount() override { return DirectX::XMINT3(nx_, ny_, 0); }
    DirectX::XMFLOAT3 iGetObjectSize() override { return DirectX::XMFLOAT3(1.0f, 1.0f, 0.0f); }
    DirectX::XMFLOAT3 iGetObjectPositionOffset() override { return DirectX::XMFLOAT3(0.0f, 0.0f, 0.0f); }
    UINT iGetConstantBufferSize() override { return static_cast<UINT>(nx_) * static_cast<UINT>(ny_) * 2u; }

    void iCreateObject(std::vector<ConstantBuffer>& constantBuffer) override {
        constantBuffer.reserve(constantBuffer.size() + static_cast<size_t>(nx_) * ny_);
        for (int row = 0; row < ny_; ++row)
            for (int col = 0; col < nx_; ++col) {
                ConstantBuffer cb;
                cb.world = DXViewer::util::transformMatrix(static_cast<float>(col), static_cast<float>(row), 0.0f, 1.0f);
                cb.worldViewProj = DXViewer::util::transformMatrix(0.0f, 0.0f, 0.0f);
                cb.transInvWorld = DXViewer::util::transformMatrix(0.0f, 0.0f, 0.0f);
                cb.color = DirectX::XMFLOAT4(0.0f, 0.0f, 0.0f, 1.0f);
                constantBuffer.push_back(cb);
            }
    }

    // Object i shows the cell the reference shows for it: (x, y) = (i / n, i % n) with n = floor(sqrt(#objects)).
    void iUpdateConstantBuffer(std::vector<ConstantBuffer>& constantBuffer, int i) override {
        const int n = static_cast<int>(std::sqrt(static_cast<double>(constantBuffer.size())));
        const int x = i / n, y = i % n;
        if (x < 0 || x >= nx_ || y < 0 || y >= ny_) return;
        const uint8_t* px = &rgba_[4u * (static_cast<size_t>(x) + static_cast<size_t>(nx_) * y)];
        constantBuffer[static_cast<size_t>(i)].color =
            DirectX::XMFLOAT4(px[0] * (1.0f / 255.0f), px[1] * (1.0f / 255.0f), px[2] * (1.0f / 255.0f), 1.0f);
    }

    void iDraw(Microsoft::WRL::ComPtr<ID3D12GraphicsCommandList>& mCommandList, int, UINT, int) override {
        mCommandList->IASetPrimitiveTopology(D3D11_PRIMITIVE_TOPOLOGY_TRIANGLELIST);
        mCommandList->DrawIndexedInstanced(static_cast<UINT>(tris_.size()), 1, 0, 0, 0);
    }
    void iSetDXApp(DX12App* dxApp) override { dxapp_ = dxApp; }

    // ---- Win32 control panel hooks (src/Kobayashi.cpp:383-629) ----
    // WM_CREATE: the reference builds its buttons, labels and nine scrollbars here.  The widgets are the application's
    // business; what the simulation needs from them is the scrollbar handle of each slider, which WM_HSCROLL is matched
    // against (:557-573): the application registers them with setScrollbar(), or passes sliderHandle(i) as lParam.
    void iWMCreate(HWND, HINSTANCE) override {}
    void setScrollbar(int index, HWND h) { if (index >= 0 && index <= TEQ) sliders_[static_cast<size_t>(index)].scrollbar = h; }
    LPARAM sliderHandle(int index) const {
        const Slider& s = sliders_[static_cast<size_t>(index)];
        return s.scrollbar ? reinterpret_cast<LPARAM>(s.scrollbar) : static_cast<LPARAM>(index + 1);
    }

    void iWMCommand(HWND, UINT, WPARAM wParam, LPARAM, HINSTANCE) override {
        switch (static_cast<int>(wParam & 0xffffu)) {
            case RESET:                                   // :512-528
                parameterInit();
                if (dxapp_) dxapp_->resetSimulationState(); else resetWithoutViewer();
                break;
            case PLAY:                                    // :531-538
                updateFlag_ = !updateFlag_;
                break;
            case STOP:                                    // :540-543
                if (dxapp_) dxapp_->resetSimulationState(); else resetWithoutViewer();
                break;
            case NEXTSTEP:                                // :544-549
                iUpdate();
                if (dxapp_) { dxapp_->update(); dxapp_->draw(); }
                break;
            default: break;
        }
    }

    void iWMHScroll(HWND, WPARAM wParam, LPARAM lParam, HINSTANCE) override {
        int index = TEQ;                                  // the reference's chain of comparisons ends in `else TEQ` (:557-573)
        for (int i = 0; i < TEQ; ++i)
            if (lParam == sliderHandle(i)) { index = i; break; }
        Slider& s = sliders_[static_cast<size_t>(index)];
        switch (static_cast<unsigned>(wParam & 0xffffu)) {
            case 5u:                                      // SB_THUMBTRACK (:585-588)
                s.ivalue = static_cast<int>((wParam >> 16) & 0xffffu);
                s.value = static_cast<float>(s.ivalue) * s.ratio;
                break;
            case 0u: case 2u:                             // SB_LINELEFT, SB_PAGELEFT (:590-597)
                if (s.ivalue - s.istride >= s.imin) { s.ivalue -= s.istride; s.value -= s.stride; }
                break;
            case 1u: case 3u:                             // SB_LINERIGHT, SB_PAGERIGHT (:599-606)
                if (s.ivalue + s.istride <= s.imax) { s.ivalue += s.istride; s.value += s.stride; }
                break;
            default: break;
        }
        pushParameters();
        if (dxapp_) dxapp_->resetSimulationState(); else resetWithoutViewer();       // :616
    }
    void iWMTimer(HWND) override {}                       // the reference refreshes its two text labels (:620-624): simTime(), simFrame()
    void iWMDestory(HWND) override {}

    // what the panel shows
    float sliderValue(int index) const { return sliders_[static_cast<size_t>(index)].value; }
    int sliderPosition(int index) const { return sliders_[static_cast<size_t>(index)].ivalue; }
    int64_t simFrame() const { return sim_.simFrame(); }
    double simTime() const { return sim_.simTimeMs(); }

private:
    struct Slider {
        float value, min, max, stride;
        int ivalue, imin, imax, istride;
        float ratio;
        HWND scrollbar;
    };
    // _parameterInit (:73-96): defaults into the float members and the thumb positions
    void parameterInit() {
        static const float dv[9] = {0.0003f, 0.010f, 1.0f, 1.6f, 0.05f, 6.0f, 0.9f, 10.0f, 1.0f};
        static const int di[9] = {3, 10, 10, 16, 5, 6, 9, 10, 10};
        for (size_t i = 0; i < 9; ++i) { sliders_[i].value = dv[i]; sliders_[i].ivalue = di[i]; }
        pushParameters();
    }
    void pushParameters() {                               // the sliders ARE the reference's float members (float& in :19,24)
        kob_params p = sim_.params();
        p.tau = sliders_[TAU].value; p.epsilon_bar = sliders_[EPLSILONBAR].value; p.mu = sliders_[MU].value;
        p.K = sliders_[K].value; p.delta = sliders_[DELTA].value; p.anisotropy = sliders_[ANISOTROPY].value;
        p.alpha = sliders_[ALPHA].value; p.gamma = sliders_[GAMMA].value; p.t_eq = sliders_[TEQ].value;
        sim_.setParams(p);
    }
    void resetWithoutViewer() { std::vector<ConstantBuffer> none; iResetSimulationState(none); }
    void refreshColours() { rgba_ = sim_.renderRGBA(); }   // colour ramp on the device, one D2H copy per frame

    Kobayashi sim_;
    int nx_, ny_;
    bool updateFlag_ = true;                              // _updateFlag (src/Kobayashi.h:84)
    DX12App* dxapp_ = nullptr;
    std::vector<Slider> sliders_;
    std::vector<Vertex> quad_;
    std::vector<unsigned int> tris_;
    std::vector<uint8_t> rgba_;
};

#endif  // KOBAYASHI_SIMULATION_HPP
