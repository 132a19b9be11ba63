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
//   * Play / Stop / Next-step / Reset semantics through setUpdated(), nextStep(), iResetSimulationState().
// The Win32 control panel (iWM* hooks, :383-629) is GUI code outside the hot path: the hooks are accepted and ignored
// here; an application that wants sliders calls the parameter setters of sim() and then iResetSimulationState(),
// which is what the reference's slider handler does (:589-616).
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
    KobayashiSimulation(int x, int y, float timeStep, int device = 0)
        : sim_(x, y, timeStep, KOB_F32, KOB_KERNEL_FAST, device), nx_(x), ny_(y) {
        quad_ = {Vertex{DirectX::XMFLOAT3(-0.5f, -0.5f, 0.0f)}, Vertex{DirectX::XMFLOAT3(-0.5f, +0.5f, 0.0f)},
                 Vertex{DirectX::XMFLOAT3(+0.5f, +0.5f, 0.0f)}, Vertex{DirectX::XMFLOAT3(+0.5f, -0.5f, 0.0f)}};
        tris_ = {0u, 1u, 2u, 0u, 2u, 3u};
        refreshColours();
    }

    Kobayashi& sim() { return sim_; }

    // ---- simulation ----
    void iUpdate() override {
        sim_.iUpdate();
        refreshColours();
    }
    void iResetSimulationState(std::vector<ConstantBuffer>& constantBuffer) override {
        sim_.iResetSimulationState();
        refreshColours();
        for (size_t i = 0; i < constantBuffer.size(); ++i) iUpdateConstantBuffer(constantBuffer, static_cast<int>(i));
    }
    bool iIsUpdated() override { return sim_.iIsUpdated(); }
    void setUpdated(bool play) { sim_.setUpdated(play); }
    void nextStep() { sim_.nextStep(); refreshColours(); }

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

    // ---- Win32 control panel hooks: GUI only, nothing to do for the simulation ----
    void iWMCreate(HWND, HINSTANCE) override {}
    void iWMCommand(HWND, UINT, WPARAM, LPARAM, HINSTANCE) override {}
    void iWMHScroll(HWND, WPARAM, LPARAM, HINSTANCE) override {}
    void iWMTimer(HWND) override {}
    void iWMDestory(HWND) override {}

private:
    void refreshColours() { rgba_ = sim_.renderRGBA(); }   // colour ramp on the device, one D2H copy per frame

    Kobayashi sim_;
    int nx_, ny_;
    DX12App* dxapp_ = nullptr;
    std::vector<Vertex> quad_;
    std::vector<unsigned int> tris_;
    std::vector<uint8_t> rgba_;
};

#endif  // KOBAYASHI_SIMULATION_HPP
