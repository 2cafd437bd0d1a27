// vap_timeloop.cuh -- v2 of stage S6 (generate_motion_profile's time loop, motion_profile_generator.py:414-628).
//
// The reference's loop carries only (current_pos, current_vel) from one iteration to the next; everything else
// it appends (parameter, heading, curvature, coordinates, angular velocity) is a pure function of the position
// before the step, and the node / action-point events only insert rows and flip signs.  So the stage is split:
//   A  k_time_state     one thread per path: the exact (pos, vel) recurrence alone (3 divisions per step, each one
//                       multiplication + two residual corrections with a tabulated / hoisted reciprocal; TMA-staged rows)
//   B1 k_time_sample    one thread per (path, step): t = distance_to_time(pos), snap gathers, point evaluation,
//                       omega, and the event candidates (frac wrap of t, action-point crossings)
//   C  k_time_events    one thread per path: replay of the event logic over the candidates only (turn / wait
//                       inserts, reverse toggles, nodes_map / actions_map), the sequential `current_time += dt`
//                       chain, and the inserted rows
//   B2 k_time_finalize  one thread per (path, step): signs, heading normalisation, scatter to the final rows
// Every floating-point operation is the reference's, in the reference's order; only the schedule changed.
#pragma once
#include "vap_device.cuh"
#include "vap_velocity.cuh"

#define TS_POS 0     // pos[k]: position before main step k (pos[k+1] = position appended by step k)
#define TS_VEL 1     // current_vel after step k
#define TS_ACC 2     // accel of step k
#define TS_TV 3      // target_vel of step k
#define TS_TH 4      // snapped heading table value at t_k (before the reverse / normalisation logic)
#define TS_OM 5      // angular_vel of step k
#define TS_X 6
#define TS_Y 7
#define TS_PLANES 8

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- exact IEEE division helpers ------------------------------------------------------------------------------
// a / b for b > 0 when a == 0: the quotient is a itself (sign kept); skips the slow path of the division sequence.
// The division is fed a harmless numerator in that case: nvcc evaluates both arms of a select, and a zero numerator
// sends the whole warp through the ~90-instruction slow path of the inlined division.
__device__ __forceinline__ double div_pos(double a, double b)
{
    bool z = (a == 0.0);
    double n = z ? 1.0 : a;
    asm volatile("" : "+d"(n));          // keep nvcc from folding the select back into the division's numerator
    double q = n / b;
    return z ? a : q;
}
// a / b with r = RN(1/b) hoisted out of the loop: q0 = a*r, two fused residual corrections.  After the first
// correction the quotient is faithful, and (Markstein) one more correction with the correctly rounded reciprocal
// gives the correctly rounded quotient, i.e. exactly what '/' returns.  Tiny / huge / non-finite numerators take
// the ordinary division.  tests/test_gpu_parity.py::test_const_division_is_ieee checks 2^28 random numerators.
__device__ __forceinline__ double div_const(double a, double b, double r)
{
    double aa = fabs(a);
    if (!(aa > 1e-280 && aa < 1e280)) return div_pos(a, b);
    double q = a * r;
    q = fma(fma(-b, q, a), r, q);
    q = fma(fma(-b, q, a), r, q);
    return q;
}


// index of np.searchsorted(xs, x, side='right') - 1 on xs[i] = fl(i*dd), 32-bit arithmetic
__device__ __forceinline__ int uniform_index32(double x, double dd, double inv_dd, int D)
{
    double e = x * inv_dd;
    int k = (e > 0.0) ? ((e < (double)(D - 1)) ? __double2int_rd(e) : D - 1) : 0;
    while (k + 1 < D && (double)(k + 1) * dd <= x) k++;
    while (k >= 0 && (double)k * dd > x) k--;
    return k;
}

// (the TMA bulk-copy / mbarrier helpers live in vap_device.cuh)

// reciprocals of the lerp denominators xs[i+1] - xs[i], xs[i] = fl(i*dd) (path-independent, like the distance grid): the
// time loop divides by them with one multiplication and two residual corrections (div_const) instead of a division.
// An entry is 0 when the denominator is outside the range where that is proven exact: the step then takes the generic path.
__global__ void k_build_lerp_recip(long long n, double dd, double* __restrict__ rden)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = (double)(i + 1) * dd - (double)i * dd;
    const long long bits = __double_as_longlong(d);
    const bool safe = (d > 1e-150) && (d < 1e150) && ((bits & 0x000FFFFFFFFFFFFFLL) != 0x000FFFFFFFFFFFFFLL);
    rden[i] = safe ? 1.0 / d : 0.0;
}

// a / b given r = RN(1/b): one multiplication and two fused residual corrections (see div_const); branch-free.  The caller
// guarantees 1e-280 < |a| < 1e280 or a == 0 (a zero numerator gives a zero quotient; its sign is not kept).
__device__ __forceinline__ double div_recip(double a, double b, double r)
{
    double q = a * r;
    q = fma(fma(-b, q, a), r, q);
    q = fma(fma(-b, q, a), r, q);
    return q;
}
__device__ __forceinline__ bool recip_safe_num(double a)
{
    const double aa = fabs(a);
    return (aa < 1e280) && ((aa > 1e-280) || (a == 0.0));
}

// shared-memory load at a 32-bit shared address + constant byte offset.  `volatile` keeps the load where it is written: the
// time loop issues its five loads BEFORE the warp vote so that the vote resolves in the shadow of their latency.
template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}

// recip_safe_num on the exponent field (integer pipe): |a| in [2^-930, 2^930) -- inside (1e-280, 1e280) -- or a == 0
__device__ __forceinline__ bool recip_safe_exp(double a)
{
    const unsigned hi = (unsigned)__double2hiint(a) & 0x7fffffffu;
    return ((hi - (93u << 20)) < ((1953u - 93u) << 20)) | (a == 0.0);
}

// a / b given r = RN(1/b), SPECULATIVELY: the quotient after ONE residual correction (faithful, and almost always already
// the correctly rounded one) goes on; the second correction runs beside the dependent chain and `same` tells whether it
// changed anything.  same == true: the value returned IS div_recip's (the IEEE quotient); false: the caller redoes the step.
__device__ __forceinline__ double div_recip_spec(double a, double b, double r, bool& same)
{
    const double q0 = a * r;
    const double q1 = fma(fma(-b, q0, a), r, q0);
    const double q2 = fma(fma(-b, q1, a), r, q1);
    same = (q1 == q2);
    return q1;
}

// A: the state recurrence (motion_profile_generator.py:523,567-583).  One thread per path; the velocity row (and the lerp
// reciprocals) are staged through a per-thread shared-memory ring by TMA bulk copies one block (BLK = 128 or 64 samples) ahead, so the loop
// never waits on HBM.
//
// The kernel runs ONE warp per scheduler with a few paths per warp, so its time is the dependent-instruction latency of a
// step times the step count of the longest path (measured on B200: 8.2 cycles per dependent DADD/DMUL/DFMA, 18 per
// double<->int conversion, 29 per shared-memory load, ~100 for an inlined division whose reciprocal refinement sits on the
// chain; a taken branch costs 15-30 cycles of a lone warp, and instructions issue in order, so independent work placed
// after a stalled instruction does not start early).  Structure (round 2; 443 cycles per step on a lone path, was 729):
//   * FAST RUN: an inner loop whose body is one straight-line block -- no branch but the loop-back.  A step is computed
//     and COMMITTED unconditionally (no select on the dependent chain); `ok` collects, off the chain, every condition
//     under which the committed values are the reference's.  A lane whose step was not ok asks for the exit at the next
//     vote; outside the loop the state of before that step comes back (a shadow copy) and the step is redone generically.
//   * ONE warp-uniform exit per step (__any_sync): some lane's position left its staged block, or its last step was refused.
//     Everybody leaves together and the lanes that asked are served, so no lane ever waits at a reconvergence point for the
//     other lanes' events (a per-lane `break` made the first lane out wait for ALL the others: +40 % time).  The vote is
//     taken on the position itself (pos >= xs[lo + 128], resolved long before the integer index) and completes while the
//     address is computed and the five loads are issued, so the branch behind them does not wait.
//   * a ring slot holds a BLK-sample block PLUS the two samples behind it (the copy is BLK + 2 doubles), so the three samples
//     and two reciprocals of a step sit at fixed offsets from one address whatever the position inside the block.
//   * floor(pos / dd) comes from one round-toward-zero addition of 2^52 (the integer is the low word of the sum and the
//     double is the sum minus 2^52), not from F2I + FRND.
//   * the three quotients go on after ONE residual correction; the second correction runs beside the chain and the step is
//     only ok if it changed nothing (div_recip_spec), i.e. if the value used IS the IEEE quotient.
//   * max(tvm, 0.001) is taken as tvm (ok only if that is what max returns); the clips pick among candidates that are
//     already there: both comparisons of a pair are issued together, -max_dec dt and max_acc dt do not wait for the division.
//   * block change (service code, outside the loop): when the position enters the next block, the block behind it is
//     refilled with the block after next -- issued a whole block (about twenty steps) before its first use.
//   * GENERIC STEP (service code: first / last intervals of a path, odd operands; a handful per path): the reference's
//     step with ordinary divisions and global loads.  It is always correct, so every service makes progress.
// Both produce the reference's bits: the fast run's quotients are verified IEEE quotients, its index is verified against
// xs[i] = fl(i dd) exactly as np.searchsorted defines it, and the candidate selects reproduce np.clip's order.
// Tried and measured on cfg4 (one path, 263 228 steps; time stage 96.4 ms with the round-1 kernel): per-lane break 69 ms
// (but +8 % on batches, see above); vote + predicated commit 74; + speculative quotients 68.6; + tvm speculation 66.4;
// + unconditional commit and the vote on the position 59.3; |da| range test on the fp pipe instead of the exponent field
// 57.9 (kept); also speculating v > 0.1 and vn > 0: 87 (too many generic steps on a path with 800 nodes); integer tests
// of the reciprocals' high words instead of r1 r2 > 0: 71; #pragma unroll 2: 62.
// BLK samples per staged block: 128 for batches whose rings all fit on the chip at once (fewest block changes), 64 for larger
// ones (half the shared memory per path: the other batch's kernels in flight keep theirs; the 2^20-path job runs 4.6 %
// faster).  Per thread: two velocity slots, then two reciprocal slots of BLK + 4 doubles (the block, the 2 samples behind it,
// 2 of padding: 16-byte multiples).
__host__ __device__ constexpr int ts_slot(int BLK) { return BLK + 4; }
__host__ __device__ constexpr int ts_stride(int BLK) { return 4 * ts_slot(BLK); }
template <int BLK>
__global__ void __launch_bounds__(32) k_time_state(long long B, const double* __restrict__ cons,
                                                   const int* __restrict__ status, double dt, double dd,
                                                   const double* __restrict__ total_len, long long D_cap,
                                                   const int* __restrict__ n_samples, const double* __restrict__ vel,
                                                   long long M_cap, double* __restrict__ stage, int* __restrict__ n_main,
                                                   const double* __restrict__ rden, long long n_rden)
{
    extern __shared__ __align__(16) double s_ring[];
    // this thread's two mbarriers (one per ring slot) live behind the rings
    const unsigned mb = (unsigned)__cvta_generic_to_shared(s_ring + (size_t)blockDim.x * ts_stride(BLK) + 2 * threadIdx.x);
    mbar_init(mb, 1);
    mbar_init(mb + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // the lanes of a warp leave the fast run TOGETHER (a vote), so every lane with a path stays in the loop until the warp
    // is done; a failed path takes part in the votes as "finished" (its own stage row takes the stores nobody reads)
    if (b >= B) return;
    const long long left = B - (long long)blockIdx.x * blockDim.x;                    // paths from this CTA's first one on
    const int nl = left < (long long)blockDim.x ? (int)left : (int)blockDim.x;       // lanes of this CTA that have a path
    const unsigned mask = nl >= 32 ? 0xffffffffu : ((1u << nl) - 1u);
    const bool part = status[b] == ST_OK;
    if (!part) n_main[b] = 0;
    const double L = part ? total_len[b] : 0.0;
    const double max_acc = part ? cons[b * 6 + 1] : 1.0, max_dec = part ? cons[b * 6 + 2] : 1.0;
    const int D = part ? n_samples[b] : 4;
    const double* vv = vel + (size_t)b * D_cap;
    const double inv_dd = 1.0 / dd, inv_dt = 1.0 / dt;
    const size_t plane = (size_t)B * (M_cap + 1);
    double* P = stage + TS_POS * plane + (size_t)b * (M_cap + 1);
    const double* ring = s_ring + (size_t)threadIdx.x * ts_stride(BLK);
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    if (rden == nullptr) n_rden = 0;                   // no table: the fast run is disabled
    const int nblk = (int)((D_cap + BLK - 1) / BLK);
    unsigned phase = 0, pend = 0;                      // per slot: parity to wait for / a copy is in flight
    auto wait_slot = [&](int sl) {
        if ((pend >> sl) & 1u) {
            mbar_wait(mb + 8 * sl, (phase >> sl) & 1u);
            phase ^= 1u << sl;
            pend &= ~(1u << sl);
        }
    };
    auto stage_block = [&](int blk) {          // rows are padded to a multiple of BLK samples by the host
        const int sl = blk & 1;
        wait_slot(sl);                          // never two copies in flight on one barrier
        if (blk < nblk) {
            // the block and the two samples behind it (they belong to the next block of the same row; the row's last
            // block has nothing behind it and the fast run never reads there: i1 + 2 <= D - 2)
            const unsigned nv = (blk + 1 < nblk) ? BLK + 2 : BLK;
            const long long rem = n_rden - (long long)blk * BLK;        // reciprocals left from this block on
            const unsigned nr_ = rem >= BLK + 2 ? BLK + 2 : (rem >= BLK ? BLK : 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier reads of the slot precede the TMA writes
            mbar_expect_tx(mb + 8 * sl, (nv + nr_) * (unsigned)sizeof(double));
            VAP_CHECK(10, (size_t)blk * BLK + nv <= (size_t)D_cap && nv <= ts_slot(BLK) && nr_ <= ts_slot(BLK) &&
                              (long long)blk * BLK + nr_ <= n_rden);
            bulk_g2s(ring_s + sl * (unsigned)(ts_slot(BLK) * sizeof(double)), vv + (size_t)blk * BLK, nv * (unsigned)sizeof(double), mb + 8 * sl);
            if (nr_) bulk_g2s(ring_s + (2 + sl) * (unsigned)(ts_slot(BLK) * sizeof(double)), rden + (size_t)blk * BLK,
                              nr_ * (unsigned)sizeof(double), mb + 8 * sl);
            pend |= 1u << sl;
        }
    };
    int blk_lo = 0;                             // block blk_lo is resident and complete; block blk_lo+1 is in flight or resident
    if (part) {
        stage_block(0);
        stage_block(1);
        wait_slot(0);
    }
    double pos = 0.0, v = part ? vv[0] : 0.0;
    const double vlast = part ? vv[D - 1] : 0.0;
    int k = 0;                                          // 32-bit: the row limit is far below 2^31
    const double hdt = 0.1 * dt;
    // fast-run index range: i1 <= D-4 (both lerps strictly inside the row) and i1 + 1 inside the staged reciprocals
    const long long nr = (n_rden / BLK) * BLK;    // the reciprocals are staged in whole blocks
    const long long il = ((long long)D - 3 < nr - 2) ? (long long)D - 3 : nr - 2;
    const double dlim = (double)(il > 0 ? il : 0);     // fast run: 0 <= pos / dd < dlim  (dlim < 2^31)
    const double ndec = -max_dec;
    const double lo_dt = ndec * dt, hi_dt = max_acc * dt;
    long long k_lim64 = 16 * M_cap + 1000000;           // far beyond any terminating profile of this capacity class
    if (k_lim64 > VAP_ROW_LIMIT) k_lim64 = VAP_ROW_LIMIT;
    const int k_limit = (int)k_lim64;
    const int m_cap = (int)(M_cap < 2147483647LL ? M_cap : 2147483647LL);
    // the fast run stores without a clamp, and its clip candidates assume -max_dec < max_acc (np.clip(x, lo, hi) of an
    // x <= lo is lo only then); otherwise every step is a generic step
    const int k_fast = (ndec < max_acc) ? (k_limit < m_cap ? k_limit : m_cap) : 0;
    const double two52 = 4503599627370496.0;
    bool fin = !part;                                   // this lane's path is finished (or failed before this stage)
    double pos_fin = 0.0;                               // its final position (a finished lane keeps stepping on scratch values)
    bool okp = true;                                    // the lane's last fast step was committed
    bool req = false;                                   // this lane asked for the exit from the fast run
    while (__any_sync(mask, !fin)) {                    // the fast run is left by a request only: never enter it without a live lane
        // ---------------- fast run over block blk_lo ----------------
        {
            const int lo = blk_lo * BLK;
            // shared address of the block: [rb + 8 (i - lo)] = vel[i], [rb + 8 (2 ts_slot(BLK) + i - lo)] = rden[i]
            const unsigned rb = ring_s + (unsigned)(blk_lo & 1) * (unsigned)(ts_slot(BLK) * sizeof(double));
            double* sp = P + (k < 0 ? 0 : (k < m_cap ? k : m_cap));   // stage row of step k (the other planes follow at `plane`)
            const double pos_hi = (double)(lo + BLK) * dd;     // xs[lo + BLK]: from here on the position is in the next block
            double pos_s = pos, v_s = v;                          // the state before the last commit
            for (;;) {
                const double e = pos * inv_dd;
                const double em = __dadd_rz(e, two52);            // 2^52 + floor(e) for 0 <= e < 2^31
                // ONE warp-uniform exit per step: some lane's position left its block, or its last step was refused.  The
                // test is on the position itself (resolved long before the integer index), and the vote completes while
                // the address is computed and the five loads are issued: the branch behind them does not wait.
                req = ((pos >= pos_hi) | !okp) & !fin;
                const bool leave = __any_sync(mask, req);
                const int i1 = __double2loint(em);
                const double ef = em - two52;                     // == trunc(e) there
                const unsigned off = (unsigned)(i1 - lo);
                const unsigned q = rb + 8u * (off < (unsigned)(BLK - 1) ? off : (unsigned)(BLK - 1));   // memory-safe whatever pos is
                VAP_CHECK(11, q >= ring_s && q + (2 * ts_slot(BLK) + 2) * 8 <= ring_s + ts_stride(BLK) * 8 && (q & 7u) == 0);
                const double y0 = lds_f64<0>(q), y1 = lds_f64<8>(q), y2 = lds_f64<16>(q);
                const double r1 = lds_f64<2 * ts_slot(BLK) * 8>(q), r2 = lds_f64<2 * ts_slot(BLK) * 8 + 8>(q);
                if (leave) break;
                const double x2 = pos + dd;
                const double x0 = ef * dd, x1 = (ef + 1.0) * dd, xx2 = (ef + 2.0) * dd;      // (double)(i1 + j) == ef + j exactly
                const double n1 = (pos - x0) * (y1 - y0), n2 = (x2 - x1) * (y2 - y1);
                bool s1, s2, s3;
                const double tv1 = y0 + div_recip_spec(n1, x1 - x0, r1, s1);
                const double tv2 = y1 + div_recip_spec(n2, xx2 - x1, r2, s2);
                const double tvm = (tv1 + tv2) / 2;
                const bool s0 = !(0.001 > tvm);                   // max(tvm, 0.001) == tvm, or the step is redone generically
                const double tv = tvm;
                const double da = tv - v;
                // da == +0 gives +0 (da is never -0: tv >= 0.001).  |da| needs no lower bound: v >= 0 and tv >= 0.001 (s0), so
                // the difference is 0, or exact and >= ulp(0.0005) (Sterbenz), or >= tv / 2; `ok` tests the upper bound / NaN.
                const double ar = div_recip_spec(da, dt, inv_dt, s3);
                // np.clip(ar, -max_dec, max_acc) and accel dt: both comparisons at once, then selects among ready values
                const bool c1 = ar > ndec, pb = ar < max_acc;
                const double prod = ar * dt;
                const double accel = c1 ? (pb ? ar : max_acc) : ndec;
                const double inc = c1 ? (pb ? prod : hi_dt) : lo_dt;
                const double vn = v + inc;
                // np.clip(vn, 0, tv)
                const double zsel = (0.0 < tv) ? 0.0 : tv;
                const bool c3 = vn > 0.0, p4 = vn < tv;
                const double v_new = c3 ? (p4 ? vn : tv) : zsel;
                const double half = 0.5 * accel * dt * dt;
                const double dpos = ((v_new <= 0.1) ? hdt : v_new * dt) + half;
                const double pos_new = pos + dpos;
                // !fin: a finished lane keeps stepping on scratch values while the rest of its warp works; with steps that move
                // backwards those values can wander back below L into a perfectly valid-looking state, and the lane must
                // not take the path up again (it did, before this test was here: n_main came out 7495 instead of 2051)
                const bool ok = !fin & (pos < L) & (e >= 0.0) & (e < dlim) & (x0 <= pos) & (pos < x1) & (x1 <= x2) & (x2 < xx2) &
                                (off < (unsigned)BLK) & (r1 * r2 > 0.0) & recip_safe_exp(n1) &
                                recip_safe_exp(n2) & (fabs(da) < 0x1p930) & (k < k_fast) & s0 & s1 & s2 & s3;
                // The commit is unconditional -- no select on the dependent chain, no branch: a refused step (and every step
                // of a finished lane) writes values nobody reads into the row's slot k, which the generic step or the final
                // store rewrites, and the lane leaves at the next vote, where the state of before the commit comes back.
                sp[0] = pos; sp[TS_VEL * plane] = v_new; sp[TS_ACC * plane] = accel; sp[TS_TV * plane] = tv;
                pos_s = pos; v_s = v;
                pos = pos_new; v = v_new;
                sp += ok; k += ok;
                okp = ok;
            }
            if (!okp) { pos = pos_s; v = v_s; }                   // undo the refused step
        }
        // ---------------- service: only the lanes that asked ----------------
        if (req) {
            if (!(pos < L)) { fin = true; pos_fin = pos; }
            else if (k >= k_limit) { k = -1; fin = true; }                 // diverging loop: report instead of hanging
            else {
                // block change: the position entered a later block
                const double e = pos * inv_dd;
                const int ic = (e > 0.0) ? ((e < 2147483000.0) ? __double2int_rz(e) : 2147483000) : 0;
                bool moved = false;
                while (blk_lo + 1 < nblk && ic >= (blk_lo + 1) * BLK) { blk_lo++; stage_block(blk_lo + 1); moved = true; }
                if (moved) wait_slot(blk_lo & 1);
                else {
                    // generic step (first / last intervals of a path, odd operands, positions outside the staged block):
                    // always correct, so every service makes progress (a block change or a step).
                    // rows have M_cap + 1 slots: steps beyond the capacity (the path is then re-run with a larger one) all
                    // land in the last slot, so the stores need no predicate
                    const int ks = k < m_cap ? k : m_cap;
                    P[ks] = pos;
                    const double x2 = pos + dd;
                    double tv1, tv2;
                    const int j1 = uniform_index32(pos, dd, inv_dd, D);
                    const int j2 = uniform_index32(x2, dd, inv_dd, D);
                    if (j1 < 0) tv1 = vv[0];
                    else if (j1 >= D - 1) tv1 = vlast;
                    else {
                        double a0 = (double)j1 * dd, a1 = (double)(j1 + 1) * dd, b0 = vv[j1], b1 = vv[j1 + 1];
                        tv1 = b0 + div_pos((pos - a0) * (b1 - b0), a1 - a0);
                    }
                    if (j2 < 0) tv2 = vv[0];
                    else if (j2 >= D - 1) tv2 = vlast;
                    else {
                        double a0 = (double)j2 * dd, a1 = (double)(j2 + 1) * dd, b0 = vv[j2], b1 = vv[j2 + 1];
                        tv2 = b0 + div_pos((x2 - a0) * (b1 - b0), a1 - a0);
                    }
                    const double tvm = (tv1 + tv2) / 2;
                    const double tv = (0.001 > tvm) ? 0.001 : tvm;                 // max(tvm, 0.001)
                    const double da = tv - v;
                    double accel = div_pos(da, dt);
                    accel = (accel > ndec) ? accel : ndec;                         // np.clip(accel, -max_dec, max_acc)
                    accel = (accel < max_acc) ? accel : max_acc;
                    double vn = v + accel * dt;
                    vn = (vn > 0.0) ? vn : 0.0;                                    // np.clip(v, 0, tv)
                    v = (vn < tv) ? vn : tv;
                    const double half = 0.5 * accel * dt * dt;
                    const double dpos = ((v <= 0.1) ? hdt : v * dt) + half;
                    pos += dpos;
                    P[ks + TS_VEL * plane] = v; P[ks + TS_ACC * plane] = accel; P[ks + TS_TV * plane] = tv;
                    k++;
                }
            }
        }
        okp = true;
    }
    if (!part) return;
    pos = pos_fin;
    wait_slot(0);                                // no TMA write may be in flight when the CTA's shared memory is released
    wait_slot(1);
    if (k >= 0 && k <= m_cap) P[k] = pos;
    n_main[b] = k;                                      // -1: diverged
}

// exactness check of div_const against the IEEE division (test hook)
__global__ void k_test_div_const(long long n, unsigned long long seed, double b, unsigned long long* __restrict__ bad)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // splitmix64 -> random double with a random exponent in [-40, 40) and a random sign
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    unsigned long long mant = z & 0x000FFFFFFFFFFFFFULL;
    int ex = (int)((z >> 52) % 80) - 40 + 1023;
    unsigned long long bits = ((z >> 63) << 63) | ((unsigned long long)ex << 52) | mant;
    double a = __longlong_as_double((long long)bits);
    double r = 1.0 / b;
    double q1 = div_const(a, b, r), q2 = a / b;
    if (__double_as_longlong(q1) != __double_as_longlong(q2)) atomicAdd(bad, 1ULL);
}

// B1: per-step lookups + event candidates
__global__ void __launch_bounds__(256, 8) k_time_sample(
    long long B, int N_max, int A_max, const int* __restrict__ n_nodes, const int* __restrict__ status,
    const double* __restrict__ ap_attr, const int* __restrict__ n_ap, const double* __restrict__ seg,
    const int* __restrict__ first_node, const double* __restrict__ param_end, const int* __restrict__ n_splines,
    int samples, long long Q_cap, const double* __restrict__ lut_d, const double* __restrict__ lut_t,
    const double* __restrict__ total_len, int spn, long long P_cap, const double* __restrict__ prop_k,
    const double* __restrict__ prop_h, long long M_cap, const int* __restrict__ n_main, double* __restrict__ stage,
    int* __restrict__ ev_wrap, int* __restrict__ ev_nwrap, int* __restrict__ ev_apc, int* __restrict__ ev_napc,
    const int* __restrict__ lut_inv, unsigned tiles_x)
{
    __shared__ double s_t[257];
    const PathTile pt = path_tile(tiles_x);
    long long b = pt.b;
    if (status[b] != ST_OK) return;
    long long M = n_main[b];
    if (M > M_cap || M < 0) return;            // capacity overflow (the caller re-runs with a larger M_cap) or diverged
    long long k0 = (long long)pt.x * blockDim.x;
    if (k0 >= M) return;
    long long k = k0 + threadIdx.x;
    const int n = n_nodes[b];
    const double* ld = lut_d + (size_t)b * Q_cap;
    const double* lt = lut_t + (size_t)b * Q_cap;
    PathGeo g = path_geo(b, N_max, seg, first_node, param_end, n_splines);
    const int Q = samples * g.S;
    const double L = total_len[b];
    const size_t plane = (size_t)B * (M_cap + 1);
    const size_t row = (size_t)b * (M_cap + 1);
    const PropGrid pg = prop_grid(spn, n);
    const int* inv = lut_inv ? lut_inv + (size_t)b * (Q_cap + LUT_INV_HDR + 2) : nullptr;
    auto d2t = [&](double d) { return inv ? distance_to_time_inv(ld, lt, inv, Q, L, n, d) : distance_to_time32(ld, lt, Q, L, n, d); };
    double t = 0.0;
    if (k < M) {
        double pos = stage[TS_POS * plane + row + k];
        t = d2t(pos);
        double curvature, heading, cx, cy;
        snap_gather2_32(prop_k + (size_t)b * P_cap, prop_h + (size_t)b * P_cap, t, pg, curvature, heading);
        eval_path<0>(g, t, cx, cy);
        double tv = stage[TS_TV * plane + row + k];
        stage[TS_TH * plane + row + k] = heading;
        stage[TS_OM * plane + row + k] = tv * curvature * -1;
        stage[TS_X * plane + row + k] = cx;
        stage[TS_Y * plane + row + k] = cy;
    }
    s_t[threadIdx.x + 1] = t;
    if (threadIdx.x == 0) s_t[0] = (k0 > 0) ? d2t(stage[TS_POS * plane + row + k0 - 1]) : 0.0;
    __syncthreads();
    if (k >= M) return;
    double tp = s_t[threadIdx.x];
    if (frac1(t) < frac1(tp) && t < (double)(n - 1)) {      // :527
        int slot = atomicAdd(ev_nwrap + b, 1);
        if (slot < N_max) ev_wrap[(size_t)b * N_max + slot] = (int)k;
    }
    int A = n_ap ? n_ap[b] : 0;
    for (int j = 0; j < A; j++) {
        double x = ap_attr[((size_t)b * A_max + j) * APA + P_T];
        if (tp < x && x < t) {                               // :549 (strict on both sides)
            int slot = atomicAdd(ev_napc + (size_t)b * A_max + j, 1);
            if (slot < EV_AP_CAND) ev_apc[((size_t)b * A_max + j) * EV_AP_CAND + slot] = (int)k;
        }
    }
}

__device__ __forceinline__ double final_heading(double th_raw, bool rev)
{   // :559-563
    double h = th_raw - (rev ? VAP_PI : 0.0);
    h = pymod_pos(h + VAP_PI, 2 * VAP_PI) - VAP_PI;
    return h * -1;
}

// C: events, inserted rows, the current_time chain, maps.  seg_k/seg_off/seg_rev[B][E_cap]: for main steps
// k >= seg_k[j] the output row is k + seg_off[j] and the reverse state is seg_rev[j].
__global__ void __launch_bounds__(32) k_time_events(
    long long B, int N_max, int A_max, const double* __restrict__ node_attr, const int* __restrict__ node_flags,
    const int* __restrict__ n_nodes, const double* __restrict__ ap_attr, const int* __restrict__ n_ap,
    const double* __restrict__ cons, int* __restrict__ status, double dt, const double* __restrict__ seg,
    const int* __restrict__ first_node, const double* __restrict__ param_end, const int* __restrict__ n_splines,
    int spn, long long P_cap, const double* __restrict__ prop_h, const double* __restrict__ total_len,
    long long M_cap, const int* __restrict__ n_main, const double* __restrict__ stage, int* __restrict__ ev_wrap,
    const int* __restrict__ ev_nwrap, const int* __restrict__ ev_apc, const int* __restrict__ ev_napc, int E_cap,
    int* __restrict__ seg_k, int* __restrict__ seg_off, int* __restrict__ seg_rev, int* __restrict__ n_seg,
    long long T_cap, long long oplane, double* __restrict__ out, int* __restrict__ nodes_map,
    int* __restrict__ actions_map, int* __restrict__ n_maps, int* __restrict__ n_out, double* __restrict__ summary)
{
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int st = status[b];
    int* nmap = nodes_map + (size_t)b * (N_max + 1);
    int* amap = actions_map + (size_t)b * (A_max > 0 ? A_max : 1);
    int nm = 0, am = 0;
    double* sr = summary + (size_t)b * 5;
    const double L = total_len[b];
    n_seg[b] = 0;
    long long T = 0;
    double last_time = 0.0;
    if (st == ST_OK) {
        const long long M = n_main[b];
        const int n = n_nodes[b];
        const int A = n_ap ? n_ap[b] : 0;
        const double* na = node_attr + (size_t)b * N_max * NA;
        const int* nf = node_flags + (size_t)b * N_max;
        const double* apa = ap_attr + (size_t)b * A_max * APA;
        const double V = cons[b * 6 + 0], max_acc = cons[b * 6 + 1], w = cons[b * 6 + 5];
        double* o_tm = out + (size_t)b * T_cap;
        double* o_pos = o_tm + oplane; double* o_lin = o_pos + oplane; double* o_acc = o_lin + oplane;
        double* o_head = o_acc + oplane; double* o_ang = o_head + oplane; double* o_x = o_ang + oplane;
        double* o_y = o_x + oplane;
        const size_t plane = (size_t)B * (M_cap + 1);
        const size_t row = (size_t)b * (M_cap + 1);
        const double* s_pos = stage + TS_POS * plane + row;
        const double* s_th = stage + TS_TH * plane + row;
        const double* s_x = stage + TS_X * plane + row;
        const double* s_y = stage + TS_Y * plane + row;
        int* sk = seg_k + (size_t)b * E_cap;
        int* so = seg_off + (size_t)b * E_cap;
        int* sv = seg_rev + (size_t)b * E_cap;
        int ns = 0;
        if (M < 0) st = ST_DIVERGED;
        else if (M > M_cap) st = ST_CAPACITY;
        bool rev = (nf[0] & F_REVERSE) != 0;
        nmap[nm++] = 0;
        if (st == ST_OK && na[A_TURN] != 0) st = ST_INDEX;          // headings[-1] on an empty list (:440)
        long long off = 0;                 // rows inserted so far
        double time = 0.0;
        double last_pos = 0.0, last_head = 0.0, last_x = 0.0, last_y = 0.0;
#define PUSH(tm_, p_, h_, w_, x_, y_)                                                                           \
        do { long long r_ = T; if (r_ < T_cap) { o_tm[r_] = (tm_); o_pos[r_] = (p_); o_lin[r_] = 0.0; o_acc[r_] = 0.0; \
             o_head[r_] = (h_); o_ang[r_] = (w_); o_x[r_] = (x_); o_y[r_] = (y_); } T++; } while (0)
        if (st == ST_OK && na[A_WAIT] > 0) {                          // :459-476
            long long steps = (long long)(na[A_WAIT] / dt);
            if (!(na[A_WAIT] / dt < (double)VAP_ROW_LIMIT)) { st = ST_DIVERGED; steps = 0; }
            const long long P = (long long)spn * n;
            const double pstep = (double)(n - 1) / (double)(P - 1);
            double h = -1 * snap_gather(prop_h + (size_t)b * P_cap, 0.0, P, n, pstep, 1.0 / pstep);
            if (rev) h -= VAP_PI;
            if (h > VAP_PI) h -= 2 * VAP_PI;
            if (h < -VAP_PI) h += 2 * VAP_PI;
            PathGeo g = path_geo(b, N_max, seg, first_node, param_end, n_splines);
            double px, py;
            eval_path<0>(g, 0.0, px, py);
            for (long long i = 0; i < steps; i++) PUSH(time + (double)i * dt, 0.0, h, 0.0, px, py);
            time += (double)steps * dt;
            off += steps;
            last_head = h; last_x = px; last_y = py;
        }
        if (st == ST_OK) {
            sk[ns] = 0; so[ns] = (int)off; sv[ns] = rev ? 1 : 0; ns++;
            // events in step order: node crossing first, then the action point of the same step
            int* wr = ev_wrap + (size_t)b * N_max;
            int nw = ev_nwrap[b];
            if (nw > N_max) st = ST_EVENTS;
            for (int i = 1; i < nw && st == ST_OK; i++) {
                int x = wr[i], j = i - 1;
                while (j >= 0 && wr[j] > x) { wr[j + 1] = wr[j]; j--; }
                wr[j + 1] = x;
            }
            int node_idx = 0, wptr = 0, action_idx = 0, last_fire = -1;
            bool actions_dead = false;
            long long kdone = 0;            // main steps whose time stamp has been written
            while (st == ST_OK) {
                int ai = 2147483647;
                if (!actions_dead && action_idx < A) {
                    int nc = ev_napc[(size_t)b * A_max + action_idx];
                    if (nc > EV_AP_CAND) { st = ST_EVENTS; break; }
                    const int* c = ev_apc + ((size_t)b * A_max + action_idx) * EV_AP_CAND;
                    for (int q = 0; q < nc; q++) if (c[q] > last_fire && c[q] < ai) ai = c[q];
                    if (ai == 2147483647) actions_dead = true;
                }
                int wi = (wptr < nw) ? wr[wptr] : 2147483647;
                long long ke = (wi < ai) ? wi : ai;
                bool have_event = ke != 2147483647;
                long long kstop = have_event ? ke : M;
                // time stamps of main steps kdone .. kstop-1 (times.append(current_time); current_time += dt)
                for (long long k = kdone; k < kstop; k++) {
                    long long r = k + off;
                    if (r < T_cap) o_tm[r] = time;
                    last_time = time;
                    time += dt;
                }
                if (kstop > kdone) {
                    // [-1] entries of the result lists = outputs of main step kstop-1
                    long long kl = kstop - 1;
                    last_pos = s_pos[kl + 1];
                    last_head = final_heading(s_th[kl], rev);
                    last_x = s_x[kl]; last_y = s_y[kl];
                }
                kdone = kstop;
                T = kstop + off;
                if (!have_event) break;
                if (wi == ke) {
                    wptr++;
                    nmap[nm++] = (int)T;
                    node_idx += 1;
                    if (node_idx >= n) { st = ST_INDEX; break; }  // spline_manager.nodes[node_idx] (:530): IndexError
                    const double* a = na + (size_t)node_idx * NA;
                    if (a[A_TURN] != 0) {                        // handle_turn (:487-507)
                        double angle = a[A_TURN] * (VAP_PI / 180.0);
                        Trapezoid tz = trapezoid_setup(V, max_acc, fabs(angle) * w / 2, dt);
                        if (!(tz.K < VAP_ROW_LIMIT)) { st = ST_DIVERGED; break; }
                        double sgn = angle > 0 ? -1.0 : 1.0;
                        double accum = 0.0, hprev = 0.0;
                        double start_heading = last_head;
                        for (long long i = 0; i < tz.K; i++) {
                            double v = trapezoid_vel(tz, i, dt);
                            double hraw = (accum / (w / 2)) * sgn;
                            accum += v * dt;
                            double om = (i == 0) ? 0.0 : (hraw - hprev) / dt;
                            hprev = hraw;
                            double hh = hraw;
                            while (hh + start_heading > VAP_PI) hh -= 2 * VAP_PI;
                            while (hh + start_heading < -VAP_PI) hh += 2 * VAP_PI;
                            last_head = start_heading + hh;
                            PUSH(time + (double)i * dt, last_pos, last_head, om, last_x, last_y);
                        }
                        time = time + (double)tz.K * dt;
                        off += tz.K;
                    }
                    if (nf[node_idx] & F_REVERSE) rev = !rev;
                    if (a[A_WAIT] > 0) {                         // handle_wait (:509-518)
                        if (!(a[A_WAIT] / dt < (double)VAP_ROW_LIMIT)) { st = ST_DIVERGED; break; }
                        long long steps = (long long)(a[A_WAIT] / dt);
                        for (long long i = 0; i < steps; i++) PUSH(time + (double)i * dt, 0.0, last_head, 0.0, last_x, last_y);
                        if (steps > 0) last_pos = 0.0;
                        time = time + (double)steps * dt;
                        off += steps;
                    }
                }
                if (ai == ke) {
                    const double* p = apa + (size_t)action_idx * APA;
                    amap[am++] = (int)T;
                    if (p[P_WAIT] > 0) {
                        if (!(p[P_WAIT] / dt < (double)VAP_ROW_LIMIT)) { st = ST_DIVERGED; break; }
                        long long steps = (long long)(p[P_WAIT] / dt);
                        for (long long i = 0; i < steps; i++) PUSH(time + (double)i * dt, 0.0, last_head, 0.0, last_x, last_y);
                        if (steps > 0) last_pos = 0.0;
                        time = time + (double)steps * dt;
                        off += steps;
                    }
                    action_idx += 1;
                    last_fire = (int)ke;
                }
                if (ns < E_cap) { sk[ns] = (int)ke; so[ns] = (int)off; sv[ns] = rev ? 1 : 0; ns++; }
                else st = ST_EVENTS;
            }
            if (st == ST_OK) nmap[nm++] = (int)T;                 // gui/path.py:342
            if (st == ST_OK && T > T_cap) st = ST_CAPACITY;
        }
#undef PUSH
        n_seg[b] = (st == ST_OK) ? ns : 0;
        if (!(st == ST_OK || st == ST_CAPACITY)) T = 0;
        if (st == ST_CAPACITY && M > M_cap) T = M;              // lower bound of the true row count
        if (T >= VAP_ROW_LIMIT) { st = ST_DIVERGED; T = 0; }
    }
    n_out[b] = (int)T;
    n_maps[2 * b] = nm; n_maps[2 * b + 1] = am;
    status[b] = st;
    if (!(st == ST_OK || st == ST_CAPACITY)) last_time = 0.0;             // a failed path has no t_end (as in k_resample)
    sr[0] = (double)T; sr[1] = L; sr[2] = last_time; sr[3] = 0.0; sr[4] = (double)st;
}

// B2: scatter the main-loop rows (times were written by C)
__global__ void __launch_bounds__(256) k_time_finalize(long long B, const int* __restrict__ status, long long M_cap,
                                                       const int* __restrict__ n_main, const double* __restrict__ stage,
                                                       int E_cap, const int* __restrict__ seg_k,
                                                       const int* __restrict__ seg_off, const int* __restrict__ seg_rev,
                                                       const int* __restrict__ n_seg, long long T_cap,
                                                       long long oplane, double* __restrict__ out,
                                                       double* __restrict__ summary, unsigned tiles_x)
{
    __shared__ double s_max[8];
    const PathTile pt = path_tile(tiles_x);
    long long b = pt.b;
    if (status[b] != ST_OK) return;
    long long M = n_main[b];
    long long k0 = (long long)pt.x * blockDim.x;
    if (k0 >= M) return;
    long long k = k0 + threadIdx.x;
    const int ns = n_seg[b];
    const int* sk = seg_k + (size_t)b * E_cap;
    double vabs = 0.0;
    if (k < M) {
        int j = 0;
        for (int q = 1; q < ns; q++) if (sk[q] <= (int)k) j = q;
        long long r = k + seg_off[(size_t)b * E_cap + j];
        bool rev = seg_rev[(size_t)b * E_cap + j] != 0;
        const size_t plane = (size_t)B * (M_cap + 1);
        const size_t row = (size_t)b * (M_cap + 1);
        double v = stage[TS_VEL * plane + row + k];
        vabs = v;
        if (r < T_cap) {
            double* o = out + (size_t)b * T_cap + r;
            double sgn = rev ? -1.0 : 1.0;
            o[1 * oplane] = stage[TS_POS * plane + row + k + 1];
            o[2 * oplane] = v * sgn;
            o[3 * oplane] = stage[TS_ACC * plane + row + k] * sgn;
            o[4 * oplane] = final_heading(stage[TS_TH * plane + row + k], rev);
            o[5 * oplane] = stage[TS_OM * plane + row + k];
            o[6 * oplane] = stage[TS_X * plane + row + k];
            o[7 * oplane] = stage[TS_Y * plane + row + k];
        }
    }
    // max |v| for the summary row (velocities are >= 0 before the sign; atomicMax on the bit pattern is exact)
    for (int o = 16; o > 0; o >>= 1) vabs = fmax(vabs, __shfl_xor_sync(0xffffffffu, vabs, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = vabs;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) m = fmax(m, s_max[q]);
        atomicMax(reinterpret_cast<unsigned long long*>(summary + (size_t)b * 5 + 3), (unsigned long long)__double_as_longlong(m));
    }
}


// ---- output packing -----------------------------------------------------------------------------------------------
// Exclusive prefix sum of the row counts (one CTA): offsets[b] = sum_{b' < b} n_out[b'], offsets[B] = total.
__global__ void __launch_bounds__(1024) k_row_offsets(long long B, const int* __restrict__ n_out,
                                                      const int* __restrict__ status, long long T_cap,
                                                      long long* __restrict__ offsets)
{
    __shared__ long long s_part[1024];
    const int t = threadIdx.x, NT = blockDim.x;
    const long long per = (B + NT - 1) / NT;
    const long long lo = (long long)t * per, hi = (lo + per < B) ? lo + per : B;
    long long sum = 0;
    for (long long b = lo; b < hi; b++) {
        long long n = (status[b] == ST_OK) ? n_out[b] : 0;
        sum += (n < T_cap) ? n : T_cap;
    }
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        long long acc = 0;
        for (int k = 0; k < NT; k++) { long long v = s_part[k]; s_part[k] = acc; acc += v; }
        offsets[B] = acc;
    }
    __syncthreads();
    long long acc = s_part[t];
    for (long long b = lo; b < hi; b++) {
        offsets[b] = acc;
        long long n = (status[b] == ST_OK) ? n_out[b] : 0;
        acc += (n < T_cap) ? n : T_cap;
    }
}

// Gather the valid part of every row of the eight output planes into one dense block per path:
// dst[8 * offsets[b] + s * n_b + k] = out[s][b][k].  dst may be pinned host memory (the kernel then streams the
// result over PCIe itself, with no padding and no size known to the host in advance).
__global__ void __launch_bounds__(256) k_pack_rows(long long B, const int* __restrict__ n_out,
                                                   const int* __restrict__ status, long long T_cap, long long oplane,
                                                   const double* __restrict__ out,
                                                   const long long* __restrict__ offsets, double* __restrict__ dst)
{
    // Persistent, deliberately small grid: a handful of CTAs saturates PCIe, and the other tiles' kernels keep the SMs.
    // Work item = (path, 256-row block); items are dealt round-robin to the CTAs.
    const long long blocks_per_row = (T_cap + blockDim.x - 1) / blockDim.x;
    const long long items = B * blocks_per_row;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        long long b = it / blocks_per_row;
        if (status[b] != ST_OK) continue;
        long long n = n_out[b];
        if (n > T_cap) n = T_cap;
        long long k = (it - b * blocks_per_row) * blockDim.x + threadIdx.x;
        if (k >= n) continue;
        const double* src = out + (size_t)b * T_cap + k;
        double* d = dst + 8 * offsets[b] + k;
#pragma unroll
        for (int s = 0; s < 8; s++) d[(size_t)s * n] = src[(size_t)s * oplane];
    }
}


// ---- "next" row f1: numeric rows of the trajectory export (gui_manager.py:284-295) --------------------------------
// For every time sample the reference writes [0, t, x*12, y*-12, heading, v*12, omega]; dst[7*(offsets[b] + k) + c].
__global__ void __launch_bounds__(256) k_export_rows(long long B, const int* __restrict__ n_out,
                                                     const int* __restrict__ status, long long T_cap, long long oplane,
                                                     const double* __restrict__ out,
                                                     const long long* __restrict__ offsets, double* __restrict__ dst,
                                                     unsigned tiles_x)
{
    const PathTile pt = path_tile(tiles_x);
    long long b = pt.b;
    if (status[b] != ST_OK) return;
    long long n = n_out[b];
    if (n > T_cap) n = T_cap;
    long long k = (long long)pt.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* src = out + (size_t)b * T_cap + k;
    double* d = dst + 7 * (offsets[b] + k);
    d[0] = 0.0;
    d[1] = src[0];                          // times
    d[2] = src[(size_t)6 * oplane] * 12;    // coords[i][0] * 12
    d[3] = src[(size_t)7 * oplane] * -12;   // coords[i][1] * -12
    d[4] = src[(size_t)4 * oplane];         // headings
    d[5] = src[(size_t)2 * oplane] * 12;    // velocities * 12
    d[6] = src[(size_t)5 * oplane];         // angular velocities
}
