// Stand-in for DXViewer's Win32App.h, used ONLY to compile the reference translation unit
// /root/reference/src/Kobayashi.cpp unmodified, in place, on Linux (see oracle/Makefile).
// It declares just enough of the Win32 / DirectX / DXViewer surface for that file to compile; every
// GUI call is a no-op.  The only arithmetic-relevant item is PI_F, restated from
// ext/DXViewer/DXViewer-3.1.0/include/dx12header.h:22.  Nothing here is product code.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <ctime>
#include <string>
#include <vector>

constexpr float PI_F = 3.141'5926f;

// ---- Win32 scalar types and no-op window API ----
using UINT = unsigned int;
using WPARAM = std::uintptr_t;
using LPARAM = std::intptr_t;
struct HWND__ { int unused; };
using HWND = HWND__*;
using HINSTANCE = void*;
using HMENU = void*;
#ifndef NULL
#define NULL 0
#endif
#ifndef TRUE
#define TRUE 1
#endif
enum : unsigned {
    WS_CHILD = 1u, WS_VISIBLE = 2u, BS_PUSHBUTTON = 4u, SBS_HORZ = 8u, SB_CTL = 2u,
    SB_LINELEFT = 0u, SB_LINERIGHT = 1u, SB_PAGELEFT = 2u, SB_PAGERIGHT = 3u, SB_THUMBTRACK = 5u
};
inline unsigned LOWORD(WPARAM w) { return (unsigned)(w & 0xffffu); }
inline unsigned HIWORD(WPARAM w) { return (unsigned)((w >> 16) & 0xffffu); }
// every created control gets its own (fake) handle: the reference tells its scrollbars apart by HWND (src/Kobayashi.cpp:557-573)
template <typename... A> inline HWND CreateWindow(A...) { static HWND__ pool[512]; static int n = 0; return &pool[n++ % 512]; }
template <typename... A> inline int EnableWindow(A...) { return 0; }
template <typename... A> inline HWND GetDlgItem(A...) { return nullptr; }
template <typename... A> inline int SetScrollRange(A...) { return 0; }
template <typename... A> inline int SetScrollPos(A...) { return 0; }
template <typename... A> inline int SetDlgItemText(A...) { return 0; }
template <typename... A> inline int SetTimer(A...) { return 0; }
template <typename... A> inline int KillTimer(A...) { return 0; }

// ---- DirectXMath storage types ----
namespace DirectX {
struct XMFLOAT2 { float x, y; XMFLOAT2() : x(0), y(0) {} XMFLOAT2(float a, float b) : x(a), y(b) {} };
struct XMFLOAT3 { float x, y, z; XMFLOAT3() : x(0), y(0), z(0) {} XMFLOAT3(float a, float b, float c) : x(a), y(b), z(c) {} };
struct XMFLOAT4 { float x, y, z, w; XMFLOAT4() : x(0), y(0), z(0), w(0) {} XMFLOAT4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {} };
struct XMINT2 { int x, y; XMINT2() : x(0), y(0) {} XMINT2(int a, int b) : x(a), y(b) {} };
struct XMINT3 { int x, y, z; XMINT3() : x(0), y(0), z(0) {} XMINT3(int a, int b, int c) : x(a), y(b), z(c) {} };
struct XMFLOAT4X4 {
    float m[16];
    XMFLOAT4X4() : m{} {}
    XMFLOAT4X4(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8, float a9,
               float a10, float a11, float a12, float a13, float a14, float a15)
        : m{a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15} {}
};
}  // namespace DirectX

struct Vertex { DirectX::XMFLOAT3 pos; DirectX::XMFLOAT3 nor; };
struct ConstantBuffer {
    DirectX::XMFLOAT4X4 worldViewProj, world, transInvWorld;
    DirectX::XMFLOAT4 color, lightPos;
};

namespace DXViewer {
namespace util {
inline DirectX::XMFLOAT4X4 transformMatrix(float x, float y, float z, float s = 1.0f) {
    return DirectX::XMFLOAT4X4(s, 0, 0, 0, 0, s, 0, 0, 0, 0, s, 0, x, y, z, 1.0f);
}
}  // namespace util
namespace xmfloat3 {
inline DirectX::XMFLOAT3 operator+(DirectX::XMFLOAT3 a, DirectX::XMFLOAT3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline DirectX::XMFLOAT3 operator*(DirectX::XMFLOAT3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
}  // namespace xmfloat3
}  // namespace DXViewer

// ---- D3D12 command list, ComPtr ----
enum { D3D11_PRIMITIVE_TOPOLOGY_TRIANGLELIST = 4 };
struct ID3D12GraphicsCommandList {
    void IASetPrimitiveTopology(int) {}
    void DrawIndexedInstanced(UINT, UINT, UINT, int, UINT) {}
};
namespace Microsoft { namespace WRL {
template <typename T> struct ComPtr { T* p = nullptr; T* operator->() const { return p; } };
} }

// ---- DXViewer app object: the three calls Kobayashi.cpp makes, with the viewer's semantics but no Direct3D
// (ext/DXViewer/src/DX12App.cpp:124-127 resetSimulationState, :554-617 update, :619-... draw) ----
class ISimulation;
class DX12App {
public:
    void setSimulation(ISimulation* s);          // + iSetDXApp + iCreateObject, as DX12App::initialize does
    void update();                               // if (iIsUpdated()) iUpdate(); then iUpdateConstantBuffer for every object
    void draw() {}
    void resetSimulationState();                 // iResetSimulationState(_constantBuffer)
    std::vector<ConstantBuffer> _constantBuffer;
    ISimulation* _simulation = nullptr;
};

// ---- the plugin interface, same 20 virtuals as ISimulation.h:7-85, in portable C++ ----
class ISimulation {
public:
    virtual void iUpdate() = 0;
    virtual void iResetSimulationState(std::vector<ConstantBuffer>& constantBuffer) = 0;
    virtual std::vector<Vertex>& iGetVertices() = 0;
    virtual std::vector<unsigned int>& iGetIndices() = 0;
    virtual UINT iGetVertexBufferSize() = 0;
    virtual UINT iGetIndexBufferSize() = 0;
    virtual DirectX::XMINT3 iGetObjectCount() = 0;
    virtual DirectX::XMFLOAT3 iGetObjectSize() = 0;
    virtual DirectX::XMFLOAT3 iGetObjectPositionOffset() = 0;
    virtual void iCreateObject(std::vector<ConstantBuffer>& constantBuffer) = 0;
    virtual void iUpdateConstantBuffer(std::vector<ConstantBuffer>& constantBuffer, int i) = 0;
    virtual void iDraw(Microsoft::WRL::ComPtr<ID3D12GraphicsCommandList>& mCommandList, int size, UINT indexCount, int i) = 0;
    virtual void iSetDXApp(DX12App* dxApp) = 0;
    virtual UINT iGetConstantBufferSize() = 0;
    virtual bool iIsUpdated() = 0;
    virtual void iWMCreate(HWND hwnd, HINSTANCE hInstance) = 0;
    virtual void iWMCommand(HWND hwnd, UINT msg, WPARAM wParam, LPARAM lParam, HINSTANCE hInstance) = 0;
    virtual void iWMHScroll(HWND hwnd, WPARAM wParam, LPARAM lParam, HINSTANCE hInstance) = 0;
    virtual void iWMTimer(HWND hwnd) = 0;
    virtual void iWMDestory(HWND hwnd) = 0;
    virtual ~ISimulation() {}
};

inline void DX12App::setSimulation(ISimulation* s) {
    _simulation = s;
    s->iSetDXApp(this);
    _constantBuffer.clear();
    s->iCreateObject(_constantBuffer);
}
inline void DX12App::update() {
    if (!_simulation) return;
    if (_simulation->iIsUpdated()) _simulation->iUpdate();
    for (int i = 0; i < (int)_constantBuffer.size(); ++i) _simulation->iUpdateConstantBuffer(_constantBuffer, i);
}
inline void DX12App::resetSimulationState() {
    if (_simulation) _simulation->iResetSimulationState(_constantBuffer);
}
