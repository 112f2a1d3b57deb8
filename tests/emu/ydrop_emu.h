// ydrop_emu.h -- TEST INFRASTRUCTURE: a 32-lane warp emulator for mimeo_b200/csrc/ydrop_warp.cuh.
// Every lane is a fibre (ucontext); a warp collective (shuffle, ballot, reduce, syncwarp) is a rendezvous of all 32
// fibres through a double-buffered exchange array. The DPX / PRMT / funnel-shift instructions are restated in plain C
// from their PTX definitions. Lets the kernel source run on the CPU against the oracle (tests/test_ydrop_emu.py).
#pragma once
#include <limits.h>
#include <stdint.h>
#include <string.h>
#include <ucontext.h>

namespace yw {

struct EmuWarp {
    ucontext_t ctx[32], main_ctx;
    int cur = 0;
    bool finished[32];
    uint32_t xch[2][32];
    int arrived = 0;
    unsigned gen = 0;
};
extern EmuWarp* g_warp;

inline void emu_yield() {
    EmuWarp* w = g_warp;
    const int me = w->cur;
    int nx = me;
    for (int t = 0; t < 32; t++) { nx = (nx + 1) & 31; if (!w->finished[nx]) break; }
    if (nx == me) return;
    w->cur = nx;
    swapcontext(&w->ctx[me], &w->ctx[nx]);
}
// all 32 lanes deposit v; returns the array of everybody's values
inline const uint32_t* emu_exchange(uint32_t v) {
    EmuWarp* w = g_warp;
    const unsigned g = w->gen;
    w->xch[g & 1][w->cur] = v;
    if (++w->arrived == 32) { w->arrived = 0; w->gen = g + 1; }
    else while (w->gen == g) emu_yield();
    return w->xch[g & 1];
}

inline int lane_id() { return g_warp->cur; }
inline uint32_t shfl_up(uint32_t v, int d) { const int l = lane_id(); const uint32_t* a = emu_exchange(v); return l - d >= 0 ? a[l - d] : v; }
inline uint32_t shfl_down(uint32_t v, int d) { const int l = lane_id(); const uint32_t* a = emu_exchange(v); return l + d < 32 ? a[l + d] : v; }
inline uint32_t shfl(uint32_t v, int src) { const uint32_t* a = emu_exchange(v); return a[src & 31]; }
inline uint32_t ballot(bool p) { const uint32_t* a = emu_exchange(p ? 1u : 0u); uint32_t m = 0; for (int l = 0; l < 32; l++) m |= (a[l] & 1u) << l; return m; }
inline int redmax(int v) { const uint32_t* a = emu_exchange((uint32_t)v); int m = INT_MIN; for (int l = 0; l < 32; l++) m = (int)a[l] > m ? (int)a[l] : m; return m; }
inline int redmin(int v) { const uint32_t* a = emu_exchange((uint32_t)v); int m = INT_MAX; for (int l = 0; l < 32; l++) m = (int)a[l] < m ? (int)a[l] : m; return m; }
inline void syncwarp() { emu_exchange(0); }
inline uint32_t atomic_add(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }

static inline int16_t h_lo(uint32_t v) { return (int16_t)(v & 0xffffu); }
static inline int16_t h_hi(uint32_t v) { return (int16_t)(v >> 16); }
static inline uint32_t h_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | (((uint32_t)hi & 0xffffu) << 16); }
inline uint32_t vadd2(uint32_t a, uint32_t b) { return h_pack((int16_t)(h_lo(a) + h_lo(b)), (int16_t)(h_hi(a) + h_hi(b))); }   // wraps, like VIADD.16x2
inline uint32_t vmax2(uint32_t a, uint32_t b) { return h_pack(h_lo(a) > h_lo(b) ? h_lo(a) : h_lo(b), h_hi(a) > h_hi(b) ? h_hi(a) : h_hi(b)); }
inline uint32_t viaddmax2(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vadd2(a, b), c); }
inline uint32_t vmax3_2(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {          // PTX prmt.b32, default mode
    const uint64_t pool = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t n = (s >> (4 * i)) & 0xfu;
        uint32_t byte = (uint32_t)(pool >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
        r |= byte << (8 * i);
    }
    return r;
}
inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { const uint64_t v = (uint64_t)lo | ((uint64_t)hi << 32); return (uint32_t)(v >> (sh & 31u)); }
inline uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t sh) { const uint64_t v = (uint64_t)lo | ((uint64_t)hi << 32); return sh >= 32 ? hi : (uint32_t)(v >> sh); }
inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t ld8(const uint8_t* p) { return *p; }
inline void fence() {}

}  // namespace yw
