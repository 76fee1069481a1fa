// TEST HARNESS ONLY.  A cooperative emulator of one 32-lane warp for the host build of the warp-per-instance solver:
// 32 ucontext fibres run the lane code round-robin; every collective (shuffle / vote) is a rendezvous at which a lane
// deposits its operand and yields to the next lane.  Collectives must be reached in warp-uniform control flow, which
// is exactly the contract of the *_sync intrinsics they stand in for.
#pragma once
#include <ucontext.h>
#include <math.h>
#include <functional>
#include <vector>

namespace kmpc {
struct Simt {
    ucontext_t ctx[32], main;
    std::vector<char> stack[32];
    int cur = 0;
    int parity[32];
    double buf[2][32];
    std::function<void()> fn;
};
extern thread_local Simt *g_simt;

inline void simt_next() {
    Simt &s = *g_simt;
    const int me = s.cur, nx = (me + 1) & 31;
    s.cur = nx;
    swapcontext(&s.ctx[me], &s.ctx[nx]);
}
inline const double *simt_rendezvous(double v) {
    Simt &s = *g_simt;
    const int me = s.cur, ph = s.parity[me];
    s.buf[ph][me] = v;
    s.parity[me] ^= 1;
    simt_next();
    return s.buf[ph];
}
inline int w_lane() { return g_simt->cur; }
inline int w_warp() { return 0; }   // one emulated warp = one block
inline int w_warps() { return 1; }
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
// one emulated warp = one block
inline void w_block_sync() { simt_rendezvous(0.0); }
inline bool w_block_any(bool p) { return w_any(p); }
inline int w_fetch(int *queue) {  // lane 0 takes the next index, everybody learns it
    double v = 0.0;
    if (w_lane() == 0) { v = (double)*queue; *queue += 1; }
    const double *b = simt_rendezvous(v);
    return (int)b[0];
}
inline void w_count_trips(unsigned long long *total, int trips) { if (total) *total += (unsigned long long)trips; }

void simt_run(const std::function<void()> &fn);  // runs fn on 32 lanes (simt.cpp part of emul.cpp)
}  // namespace kmpc
