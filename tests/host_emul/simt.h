// TEST HARNESS ONLY.  A cooperative emulator of one thread block of W 32-lane warps for the host build of the warp-per-instance
// solver: 32 W ucontext fibres run the lane code; every warp collective (shuffle / vote) is a rendezvous at which a lane
// deposits its operand and yields to the next lane of ITS warp (the last lane wraps to the first, so a warp runs on between
// block barriers as it does on the device); a block barrier is the same rendezvous with the last lane yielding to the NEXT
// warp.  Collectives must be reached in warp-uniform control flow and block barriers by every warp, which is exactly the
// contract of the *_sync intrinsics / __syncthreads they stand in for.  Fibres never run concurrently: shared-memory atomics
// are plain read-modify-writes.
#pragma once
#include <ucontext.h>
#include <math.h>
#include <functional>
#include <vector>

namespace kmpc {
#define SIMT_MAX_WARPS 16
struct Simt {
    ucontext_t ctx[32 * SIMT_MAX_WARPS], main;
    std::vector<char> stack[32 * SIMT_MAX_WARPS];
    int cur = 0, nw = 1;
    int parity[32 * SIMT_MAX_WARPS], bparity[32 * SIMT_MAX_WARPS];   // phase of the next warp / block rendezvous of every fibre
    double buf[SIMT_MAX_WARPS][2][32];
    double bbuf[2][32 * SIMT_MAX_WARPS];
    std::function<void()> fn;
};
extern thread_local Simt *g_simt;

// yield to the next lane of this warp (block == false) or, from the last lane, to the first lane of the next warp (block == true)
inline void simt_next(bool block) {
    Simt &s = *g_simt;
    const int me = s.cur, w = me >> 5, l = me & 31;
    const int nx = l < 31 ? me + 1 : (block ? (((w + 1) % s.nw) << 5) : (w << 5));
    s.cur = nx;
    if (nx != me) swapcontext(&s.ctx[me], &s.ctx[nx]);
}
inline const double *simt_rendezvous(double v, bool block = false) {
    Simt &s = *g_simt;
    const int me = s.cur, w = me >> 5, l = me & 31, ph = s.parity[me];
    s.buf[w][ph][l] = v;
    s.parity[me] ^= 1;
    simt_next(block);
    return s.buf[w][ph];
}
inline int w_lane() { return g_simt->cur & 31; }
inline int w_warp() { return g_simt->cur >> 5; }
inline int w_warps() { return g_simt->nw; }
inline int w_block() { return 0; }
inline int w_serial_warp(int) { return 0; }
inline double w_down(double v, int d) { const int me = w_lane(); const double *b = simt_rendezvous(v); return me + d < 32 ? b[me + d] : v; }
inline double w_up(double v, int d) { const int me = w_lane(); const double *b = simt_rendezvous(v); return me - d >= 0 ? b[me - d] : v; }
inline double w_xor(double v, int m) { const int me = w_lane(); const double *b = simt_rendezvous(v); return b[me ^ m]; }
inline double w_bcast(double v, int src) { const double *b = simt_rendezvous(v); return b[src]; }
inline int w_bcast_i(int v, int src) { const double *b = simt_rendezvous((double)v); return (int)b[src]; }
inline bool w_all(bool p) { const double *b = simt_rendezvous(p ? 1.0 : 0.0); bool r = true; for (int i = 0; i < 32; ++i) r = r && b[i] != 0.0; return r; }
inline bool w_any(bool p) { const double *b = simt_rendezvous(p ? 1.0 : 0.0); bool r = false; for (int i = 0; i < 32; ++i) r = r || b[i] != 0.0; return r; }
inline void w_sync() { simt_rendezvous(0.0); }
inline unsigned w_ballot(bool p) { const double *b = simt_rendezvous(p ? 1.0 : 0.0); unsigned m = 0; for (int i = 0; i < 32; ++i) if (b[i] != 0.0) m |= 1u << i; return m; }
inline void w_reconverge(unsigned) {}  // fibres run one after the other: nothing to re-join
// emulation of the CREDUX-based reductions: max with NaN propagation / min, of sign-bit-clear doubles
inline double w_max_nn(double v) { const double *b = simt_rendezvous(v); double r = b[0]; for (int i = 1; i < 32; ++i) r = (b[i] > r || b[i] != b[i]) ? b[i] : r; return r; }
inline double w_min_nn(double v) { const double *b = simt_rendezvous(v); double r = INFINITY; bool any = false; for (int i = 0; i < 32; ++i) if (b[i] == b[i]) { any = true; r = b[i] < r ? b[i] : r; } return any ? r : NAN; }
// block barrier: every warp deposits, the last lane of each warp hands over to the next warp; the values of ALL warps are visible after it
inline const double *simt_block_rendezvous(double v) {
    Simt &s = *g_simt;
    const int me = s.cur, ph = s.bparity[me];
    s.bbuf[ph][me] = v;
    s.bparity[me] ^= 1;
    simt_next(true);
    return s.bbuf[ph];
}
inline void w_block_sync() { simt_block_rendezvous(0.0); }
inline bool w_block_any(bool p) {
    const double *b = simt_block_rendezvous(p ? 1.0 : 0.0);
    bool r = false;
    for (int i = 0; i < 32 * g_simt->nw; ++i) r = r || b[i] != 0.0;
    return r;
}
inline unsigned w_smem_or(unsigned *p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
inline int w_block_warps_with(bool p) {
    const double *b = simt_block_rendezvous(p ? 1.0 : 0.0);
    int n = 0;
    for (int w = 0; w < g_simt->nw; ++w) n += b[32 * w] != 0.0 ? 1 : 0;
    return n;
}
inline int w_fetch(int *queue) {  // lane 0 takes the next index, everybody learns it
    double v = 0.0;
    if (w_lane() == 0) { v = (double)*queue; *queue += 1; }
    const double *b = simt_rendezvous(v);
    return (int)b[0];
}
inline void w_count_trips(unsigned long long *total, int trips) { if (total) *total += (unsigned long long)trips; }
inline int w_take_slot(int *counter) { return (*counter)++; }

void simt_run(const std::function<void()> &fn, int warps = 1);  // runs fn on 32 * warps lanes (emul.cpp)
}  // namespace kmpc
