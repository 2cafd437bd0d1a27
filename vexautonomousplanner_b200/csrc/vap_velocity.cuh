// vap_velocity.cuh -- stages S3-events / S4 / S5: sample-parallel event detection, a pre-pass that hoists every term of
// the forward / backward recurrences that depends neither on the velocity state nor on the events into five values per
// sample (chunk-interleaved rows, written once), and chunk-speculative kernels that run the exact serial recurrences on
// many chunks of one path at once over coalesced streams.
//
// Exactness: the recurrences of motion_profile_generator.py:188-311 are NOT associative (the wheel-acceleration
// term depends on v[i] and v[i-1]), so no scan is used.  A chunk starts from a guessed state, and
// is re-run from its predecessor's true end state until two consecutive velocities are BITWISE equal to the ones
// computed before; from there on the old results are the serial results.  The fix-up loop ends when no chunk
// changed, at which point (by induction from chunk 0) every value equals the serial evaluation bit for bit.
#pragma once
#include "vap_device.cuh"

#define EV_AP_CAND 4

// read-only (ld.global.nc) load of a 32-byte record
__device__ __forceinline__ double4 ldg_d4(const double4* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// ---- S3 (parallel): t, kappa, theta per distance sample + event candidates -----------------------------------
// wrap candidates: samples with frac(t[i-1]) > frac(t[i]) and t[i] < N-1  (motion_profile_generator.py:124)
// action candidates: samples with t[i-1] < ap.t <= t[i]                   (:142-146)
__global__ void __launch_bounds__(256, 8) k_dist_sample_ev(
    int N_max, int A_max, const int* __restrict__ n_nodes, const int* __restrict__ n_splines,
    const int* __restrict__ status, const double* __restrict__ ap_attr, const int* __restrict__ n_ap,
    const double* __restrict__ dgrid, int samples, long long Q_cap, const double* __restrict__ lut_d,
    const double* __restrict__ lut_t, const double* __restrict__ total_len, int spn, long long P_cap,
    const double* __restrict__ prop_k, const double* __restrict__ prop_h, long long D_cap,
    const int* __restrict__ n_samples, double* __restrict__ t_out, double* __restrict__ kap, double* __restrict__ th,
    int* __restrict__ ev_wrap, int* __restrict__ ev_nwrap, int* __restrict__ ev_apc, int* __restrict__ ev_napc,
    const int* __restrict__ lut_inv, unsigned tiles_x)
{
    __shared__ double s_t[257];
    const PathTile pt = path_tile(tiles_x);
    long long b = pt.b;
    if (status[b] != ST_OK) return;
    const int D = n_samples[b];
    const int i0 = pt.x * blockDim.x;
    if (i0 >= D) return;
    const int i = i0 + threadIdx.x;
    const int n = n_nodes[b];
    const double* ld = lut_d + (size_t)b * Q_cap;
    const double* lt = lut_t + (size_t)b * Q_cap;
    const int Q = samples * n_splines[b];
    const double L = total_len[b];
    const PropGrid pg = prop_grid(spn, n);
    const int* inv = lut_inv ? lut_inv + (size_t)b * (Q_cap + LUT_INV_HDR + 2) : nullptr;
    auto d2t = [&](double d) { return inv ? distance_to_time_inv(ld, lt, inv, Q, L, n, d) : distance_to_time32(ld, lt, Q, L, n, d); };
    double t = 0.0;
    if (i < D) {
        t = (i == D - 1) ? (double)(n - 1) : d2t(dgrid[i]);
        double k, h;
        snap_gather2_32(prop_k + (size_t)b * P_cap, prop_h + (size_t)b * P_cap, t, pg, k, h);
        size_t o = (size_t)b * D_cap + i;
        if (t_out) t_out[o] = t;           // the parameters themselves are only an inspection output
        kap[o] = k; th[o] = h;
    }
    s_t[threadIdx.x + 1] = t;
    if (threadIdx.x == 0) s_t[0] = (i0 > 0) ? d2t(dgrid[i0 - 1]) : 0.0;   // prev_t of sample 0 is 0
    __syncthreads();
    if (i >= D - 1) return;            // the final appended sample takes no part in the event logic
    double tp = s_t[threadIdx.x];
    if (frac1(tp) > frac1(t) && t < (double)(n - 1)) {
        int slot = atomicAdd(ev_nwrap + b, 1);
        if (slot < N_max) ev_wrap[(size_t)b * N_max + slot] = (int)i;
    }
    int A = n_ap ? n_ap[b] : 0;
    for (int k = 0; k < A; k++) {
        double x = ap_attr[((size_t)b * A_max + k) * APA + P_T];
        if (tp < x && t >= x) {
            int slot = atomicAdd(ev_napc + (size_t)b * A_max + k, 1);
            if (slot < EV_AP_CAND) ev_apc[((size_t)b * A_max + k) * EV_AP_CAND + slot] = (int)i;
        }
    }
}

// ---- S3 (per path): replay the sampling loop's event logic over the candidate samples only --------------------
// Outputs the reference's max_accels / boundary_map plus the piecewise-constant initial-velocity regimes:
//   vr_idx/vr_val: v0[i] = vr_val[j] for the last j with vr_idx[j] <= i      (max_velocity before sample i's events)
//   st_idx       : samples whose initial velocity is overwritten with 0.01    (stop nodes / stop action points)
__global__ void k_resolve_events(long long B, int N_max, int A_max, const double* __restrict__ node_attr,
                                 const int* __restrict__ node_flags, const int* __restrict__ n_nodes,
                                 const double* __restrict__ ap_attr, const int* __restrict__ ap_flags,
                                 const int* __restrict__ n_ap, const double* __restrict__ cons,
                                 int* __restrict__ status, int* __restrict__ ev_wrap, const int* __restrict__ ev_nwrap,
                                 const int* __restrict__ ev_apc, const int* __restrict__ ev_napc, int E_cap,
                                 double* __restrict__ max_accels, int* __restrict__ bidx, int* __restrict__ bval,
                                 int* __restrict__ n_ev, int* __restrict__ vr_idx, double* __restrict__ vr_val,
                                 int* __restrict__ st_idx, int* __restrict__ n_vr, double dt,
                                 float* __restrict__ ins_est)
{
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    n_ev[2 * b] = 0; n_ev[2 * b + 1] = 0; n_vr[2 * b] = 0; n_vr[2 * b + 1] = 0;
    if (ins_est) ins_est[b] = 0.f;
    if (status[b] != ST_OK) return;
    const int n = n_nodes[b];
    const int A = n_ap ? n_ap[b] : 0;
    const double* na = node_attr + (size_t)b * N_max * NA;
    const int* nf = node_flags + (size_t)b * N_max;
    const double* apa = ap_attr + (size_t)b * A_max * APA;
    const int* apf = ap_flags + (size_t)b * A_max;
    const double V = cons[b * 6 + 0], A0 = cons[b * 6 + 1];
    double* ma = max_accels + (size_t)b * E_cap;
    int* bi = bidx + (size_t)b * E_cap;
    int* bv = bval + (size_t)b * E_cap;
    int* vi = vr_idx + (size_t)b * E_cap;
    double* vv = vr_val + (size_t)b * E_cap;
    int* si = st_idx + (size_t)b * E_cap;
    int* wr = ev_wrap + (size_t)b * N_max;
    int nw = ev_nwrap[b];
    if (nw > N_max) { status[b] = ST_EVENTS; return; }
    for (int i = 1; i < nw; i++) {           // sort wrap samples ascending
        int x = wr[i], j = i - 1;
        while (j >= 0 && wr[j] > x) { wr[j + 1] = wr[j]; j--; }
        wr[j + 1] = x;
    }
    int n_acc = 0, n_b = 0, nvr = 0, nst = 0;
    double max_velocity = (na[A_MAXVEL] > 0) ? na[A_MAXVEL] : V;
    ma[n_acc++] = (na[A_MAXACC] > 0) ? na[A_MAXACC] : A0;
    bi[0] = 0; bv[0] = 0; n_b = 1;
    vi[nvr] = 0; vv[nvr] = max_velocity; nvr++;
    // action point k fires at its smallest candidate sample after the previous action point's sample
    int node_num = 0, wptr = 0, action_idx = 0, last_fire = -1;
    bool actions_dead = false;
    while (true) {
        // next action-point sample (if any)
        int ai = 2147483647;
        if (!actions_dead && action_idx < A) {
            int nc = ev_napc[(size_t)b * A_max + action_idx];
            if (nc > EV_AP_CAND) { status[b] = ST_EVENTS; return; }
            const int* c = ev_apc + ((size_t)b * A_max + action_idx) * EV_AP_CAND;
            for (int k = 0; k < nc; k++) if (c[k] > last_fire && c[k] < ai) ai = c[k];
            if (ai == 2147483647) actions_dead = true;      // this action point never fires -> none after it does
        }
        int wi = (wptr < nw) ? wr[wptr] : 2147483647;
        if (wi == 2147483647 && ai == 2147483647) break;
        int i = wi < ai ? wi : ai;
        bool stop = false;
        if (wi == i) {                       // node crossing first (:124-140)
            wptr++;
            node_num += 1;
            const double* a = na + (size_t)node_num * NA;
            if (nf[node_num] & F_STOP) stop = true;
            max_velocity = (a[A_MAXVEL] > 0) ? a[A_MAXVEL] : V;
            ma[n_acc++] = (a[A_MAXACC] > 0) ? a[A_MAXACC] : A0;
            if (node_num < n - 1) {
                if (bi[n_b - 1] == i) bv[n_b - 1] = n_acc - 1;
                else { bi[n_b] = i; bv[n_b] = n_acc - 1; n_b++; }
            }
        }
        if (ai == i) {                       // then the action point (:142-163)
            const double* p = apa + (size_t)action_idx * APA;
            max_velocity = (p[P_MAXVEL] > 0) ? p[P_MAXVEL] : V;
            if (apf[action_idx] & F_STOP) stop = true;
            ma[n_acc++] = (p[P_MAXACC] > 0) ? p[P_MAXACC] : A0;
            if (bi[n_b - 1] == i) bv[n_b - 1] = n_acc - 1;
            else { bi[n_b] = i; bv[n_b] = n_acc - 1; n_b++; }
            action_idx += 1;
            last_fire = i;
        }
        if (stop) si[nst++] = i;
        if (vi[nvr - 1] == i + 1) vv[nvr - 1] = max_velocity;
        else { vi[nvr] = i + 1; vv[nvr] = max_velocity; nvr++; }
    }
    ma[n_acc++] = A0;
    n_ev[2 * b] = n_acc; n_ev[2 * b + 1] = n_b;
    n_vr[2 * b] = nvr; n_vr[2 * b + 1] = nst;
    if (ins_est) {
        // rows the time-domain stage inserts for waits and turn profiles (upper bound: assume every one fires)
        const double w = cons[b * 6 + 5];
        double extra = 0.0;
        for (int i = 0; i < n; i++) {
            const double* a = na + (size_t)i * NA;
            if (a[A_WAIT] > 0) extra += floor(a[A_WAIT] / dt);
            if (a[A_TURN] != 0) {
                double angle = a[A_TURN] * (VAP_PI / 180.0);
                Trapezoid tz = trapezoid_setup(V, A0, fabs(angle) * w / 2, dt);
                extra += (double)tz.K;
            }
        }
        for (int i = 0; i < A; i++) if (apa[i * APA + P_WAIT] > 0) extra += floor(apa[i * APA + P_WAIT] / dt);
        ins_est[b] = (float)extra;
    }
}

__device__ __forceinline__ double recip_for_pass(double g);   // defined with the pass kernels below

// ---- chunk-interleaved layout of everything the passes stream ---------------------------------------------------------
// A path's D-1 steps ("edges" e = 0 .. D-2: forward step e goes from sample e to e+1, backward step e+1 from sample e+1
// to e) are cut into NT chunks of Lc = ceil((D-1)/NT) edges; chunk c owns edges [c*Lc, min((c+1)*Lc, D-1)).  Thread c of a
// pass CTA walks chunk c, so at any moment the warp needs edge s of 32 different chunks: the per-edge data is stored at
//      slot(e) = (e % Lc) * NT + e / Lc          (row = position inside the chunk, column = chunk)
// which makes every warp-wide load of the passes one contiguous, fully used run of memory (1 KB of records, 256 B of
// reciprocals / forward velocities).  Rows of RS = D_cap + 256 slots per path; slot RS-1 holds the forward velocity of
// the last sample.
__device__ __forceinline__ int chunk_len(int steps, int NT) { return (steps + NT - 1) / NT; }
__device__ __forceinline__ int edge_slot(int e, int Lc, int NT) { const int c = e / Lc; return (e - c * Lc) * NT + c; }

// ---- pre-pass (parallel): everything of a pass step that depends neither on the velocity state nor on the events --------
// One row of 5 NT doubles per chunk position s: five field planes of NT columns, element (row s, field f, column c) of a
// path is rec[(5 s + f) NT + c]; the terms of the final sample D-1 sit in the last three doubles of the path's 5 RS.
// fields 0-2, per sample i (row / column of edge i):
//    { |kappa_i|, G_i, stat_i }
//       G_i    = min(vlim_i, cap_i)                         velocity caps (:212-218, :239)
//       stat_i = min(max_ang_acc/|k|, 2 A0/(w|k|+2), A0)    the state-independent acceleration limits for the path's own
//                (A0 when straight)                         max_acc A0 -- forward a_static AND backward d_static as long as
//                                                           no node / action point overrides max_acceleration
// fields 3-4, per edge e:
//    { gh_e = 2|theta_{e+1} - theta_e|, recip_for_pass(gh_e) }   denominator of the wheel-acceleration term of forward
//                                                                step e and of backward step e+1, and its reciprocal
// What depends on the events is applied by the passes themselves from small shared-memory tables: the initial velocity
// v0[i+1] (regimes, stops, end velocity) enters as C_i = min(v0[i+1], G_i).  Only a path with max_acceleration overrides
// makes the pre-pass look at the regime table: stat then uses the forward regime's max_acc and the backward pass's limit
// (its max_dec) goes to a slot-order array of its own, statB.
// One thread per SLOT (coalesced stores); kappa / theta come through a shared-memory tile whose loads run along the columns.
struct SharedRecip { double b, y; bool ok; };
// a / b for several numerators over one denominator: y = RN(1/b) once, then per quotient one multiplication and two fused
// residual corrections (Markstein: with the correctly rounded reciprocal the second correction gives the correctly rounded
// quotient unless b's significand is all ones).  Operands outside the proven range take the ordinary division.
__device__ __forceinline__ SharedRecip shared_recip(double b)
{
    SharedRecip r;
    const long long bits = __double_as_longlong(b);
    r.ok = (b > 1e-150) && (b < 1e150) && ((bits & 0x000FFFFFFFFFFFFFLL) != 0x000FFFFFFFFFFFFFLL);
    r.b = b;
    r.y = 1.0 / b;
    return r;
}
__device__ __forceinline__ double div_shared(double a, const SharedRecip& r)
{
    const double aa = fabs(a);
    if (!(r.ok && aa > 1e-150 && aa < 1e150)) return a / r.b;
    double q = a * r.y;
    q = fma(fma(-r.b, q, a), r.y, q);
    q = fma(fma(-r.b, q, a), r.y, q);
    return q;
}
// does any regime of the path use a max_acc other than the path's own A0?  (max_accels[0 .. n_acc): the sampling loop's list)
__device__ __forceinline__ int pass_has_override(const double* __restrict__ ma, int n_acc, double A0)
{
    int f = 0;
    for (int j = 0; j < n_acc; j++) f |= (ma[j] != A0);
    return f;
}
__device__ __forceinline__ int last_le(const int* __restrict__ a, int n, int key)      // last j in [0, n) with a[j] <= key
{
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a[mid] <= key) lo = mid; else hi = mid - 1; }
    return lo;
}
struct SampleTerms { double ak, G, stat, stat_b; };
// acc_f: max_acc of the forward regime the sample lies in; dec_b: the backward pass's max_dec (both the path's A0 unless a
// node / action point overrides max_acceleration)
template <bool OVR>
__device__ __forceinline__ SampleTerms prepass_sample(double V, double acc_f, double dec_b, double w, double max_angular_vel,
                                                      double max_angular_accel, double k)
{
    const double ak = fabs(k);
    const bool straight = ak < 1e-6;
    // w|k| + 2 = 2 (1 + w|k|/2) bit for bit (halving and doubling are exact), so the wheel-speed cap
    // |V / (1 + w|k|/2)| (:239) IS |2V / (w|k| + 2)| (:214): one quotient serves both.
    const SharedRecip rden = shared_recip(w * ak + 2);
    const double v_kin = div_shared(2 * V, rden);
    const double cap = fabs(v_kin);
    double vlim, stat, stat_b;
    if (straight) { vlim = V; stat = acc_f; stat_b = dec_b; }
    else {
        const SharedRecip rak = shared_recip(ak);
        const double v_ang = div_shared(max_angular_vel, rak);
        const double a_ang = div_shared(max_angular_accel, rak);
        // Constraints.max_speed_at_curvature (:23-33) with 2*V/w already evaluated
        const double m = (max_angular_vel * V) / (ak * V + max_angular_vel);
        const double v_curve = pymin(m, V);
        vlim = pymin(pymin(v_ang, v_kin), v_curve);
        const double a_kin = div_shared(2 * acc_f, rden);
        stat = pymin(pymin(a_ang, a_kin), acc_f);
        stat_b = stat;
        if (OVR && dec_b != acc_f) {
            const double d_kin = div_shared(2 * dec_b, rden);
            stat_b = pymin(pymin(a_ang, d_kin), dec_b);
        }
    }
    SampleTerms t;
    t.ak = ak; t.G = pymin(vlim, cap); t.stat = stat; t.stat_b = stat_b;
    return t;
}

// per-slot work of the pre-pass for a path WITH max_acceleration overrides: the forward regime of the sample is looked up and
// the backward limit goes to its own array.  Out of line: the common (override-free) path keeps its 32 registers.
__device__ __noinline__ void prepass_slot_ovr(long long b, int e, int j, int steps, int D, int NT, long long RS, int E_cap,
                                             double V, double w, double max_angular_vel, double max_angular_accel,
                                             double k_e, double k_last, double gh,
                                             const double* __restrict__ max_accels, const int* __restrict__ bidx,
                                             const int* __restrict__ bval, const int* __restrict__ n_ev,
                                             double* __restrict__ pr, size_t o, double* __restrict__ statB)
{
    const int nb = n_ev[2 * b + 1];
    const double dec_b = max_accels[(size_t)b * E_cap + bval[(size_t)b * E_cap + nb - 1]];       // backward max_dec
    const double acc_f = max_accels[(size_t)b * E_cap + bval[(size_t)b * E_cap + last_le(bidx + (size_t)b * E_cap, nb, e)]];
    const SampleTerms t = prepass_sample<true>(V, acc_f, dec_b, w, max_angular_vel, max_angular_accel, k_e);
    pr[o] = t.ak; pr[o + NT] = t.G; pr[o + 2 * NT] = t.stat; pr[o + 3 * NT] = gh; pr[o + 4 * NT] = recip_for_pass(gh);
    statB[(size_t)b * RS + j] = t.stat_b;
    if (e == steps - 1) {
        const SampleTerms u = prepass_sample<true>(V, dec_b, dec_b, w, max_angular_vel, max_angular_accel, k_last);
        pr[5 * RS - 3] = u.ak; pr[5 * RS - 2] = u.G; pr[5 * RS - 1] = u.stat;
        statB[(size_t)b * RS + RS - 1] = u.stat_b;
    }
}

#define PP_TILES 4          // 256-slot tiles per pre-pass CTA: the per-path preamble (status, sizes, override test, constants) is paid once
__global__ void __launch_bounds__(256, 8) k_prepass(
    const int* __restrict__ status, const double* __restrict__ cons, long long D_cap, const int* __restrict__ n_samples,
    const double* __restrict__ kap, const double* __restrict__ th, int NT, long long RS, double* __restrict__ rec,
    int E_cap, const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, double* __restrict__ statB, unsigned tiles_x)
{
    extern __shared__ double s_tile[];               // two tile buffers: one barrier per tile
    const PathTile pt = path_tile(tiles_x);
    const long long b = pt.b;
    if (status[b] != ST_OK) return;
    const int D = n_samples[b];
    // does some node / action point override max_acceleration?  One regime per thread, loaded here and compared at the
    // barrier, so that the latency overlaps the tile's (paths with more than blockDim.x regimes: a loop for the rest)
    const bool o_in = (int)threadIdx.x < E_cap;
    const int o_n = n_ev[2 * b];
    const double o_ma = o_in ? max_accels[(size_t)b * E_cap + threadIdx.x] : 0.0;
    const double o_A0 = cons[b * 6 + 1];
    const int steps = D - 1;
    const int sh = 31 - __clz(NT);                   // NT is a power of two: every index split below is a shift
    const int Lc = (steps + NT - 1) >> sh;           // chunk_len(steps, NT)
    const int jend = Lc << sh;
    int j0 = pt.x * (blockDim.x * PP_TILES);
    if (steps <= 0 || j0 >= jend) return;
    // kappa / theta of a tile's slots: rows s0 .. s0+RW-1 of every column plus the row after them, fetched with the lanes
    // running ALONG a column (contiguous samples) into shared memory
    const double* kr = kap + (size_t)b * D_cap;
    const double* tr = th + (size_t)b * D_cap;
    const int rsh = 8 - sh;                          // blockDim.x == 256: RW = 256 / NT rows per tile
    const int RW = 1 << rsh;
    const int st = (RW + 1) | 1;                     // odd tile stride
    __shared__ double s_c[2];                        // per-path constants: one thread divides, not all 256
    if (threadIdx.x == 0) {
        const double V_ = cons[b * 6 + 0], A0_ = cons[b * 6 + 1], w_ = cons[b * 6 + 5];
        s_c[0] = 2 * V_ / w_;                        // max_angular_vel   (:81)
        s_c[1] = 2 * A0_ / w_;                       // max_angular_accel (:82)
    }
    int p_ovr = o_in & ((int)threadIdx.x < o_n) & (o_ma != o_A0);
    for (int q = threadIdx.x + blockDim.x; q < o_n; q += blockDim.x) p_ovr |= (max_accels[(size_t)b * E_cap + q] != o_A0);
    const double V = cons[b * 6 + 0], A0 = cons[b * 6 + 1], w = cons[b * 6 + 5];
    double* pr = rec + (size_t)b * RS * 5;
    int ovr = 0;
    for (int it = 0; it < PP_TILES && j0 < jend; ++it, j0 += blockDim.x) {
        double* t_th = s_tile + (it & 1) * (2 * NT * st);
        double* t_k = t_th + NT * st;
        const int s0 = j0 >> sh;
        {
            const int cc = threadIdx.x >> rsh, r = threadIdx.x & (RW - 1);
            int ee = cc * Lc + s0 + r;
            ee = ee > D - 1 ? D - 1 : ee;
            t_th[cc * st + r] = tr[ee];
            t_k[cc * st + r] = kr[ee];
            if (threadIdx.x < NT) {                  // the halo row
                int eh = threadIdx.x * Lc + s0 + RW;
                eh = eh > D - 1 ? D - 1 : eh;
                t_th[threadIdx.x * st + RW] = tr[eh];
            }
        }
        // the barrier of the first tile also carries the override vote and publishes s_c; the other buffer is free again
        // one barrier later, when every thread has left the tile before
        if (it == 0) ovr = __syncthreads_or(p_ovr); else __syncthreads();
        const int j = j0 + threadIdx.x;
        const int s = j >> sh, c = j & (NT - 1);
        const int e = c * Lc + s;              // this slot's edge = the sample whose terms this thread evaluates
        if (s >= Lc || e >= steps) continue;
        const double max_angular_vel = s_c[0];
        const double max_angular_accel = s_c[1];
        const int tl = c * st + (s - s0);
        const double gh = 2 * fabs(t_th[tl + 1] - t_th[tl]);
        // without overrides both passes use the path's A0 (the common case, kept free of any table code); with them the
        // forward regime of this sample is looked up and the backward limit goes to its own slot-order array
        const size_t o = (size_t)s * 5 * NT + c;
        if (ovr) {                                  // uniform over the CTA; rare
            prepass_slot_ovr(b, e, j, steps, D, NT, RS, E_cap, V, w, max_angular_vel, max_angular_accel, t_k[tl], kr[D - 1],
                             gh, max_accels, bidx, bval, n_ev, pr, o, statB);
            continue;
        }
        const SampleTerms t = prepass_sample<false>(V, A0, A0, w, max_angular_vel, max_angular_accel, t_k[tl]);
        pr[o] = t.ak; pr[o + NT] = t.G; pr[o + 2 * NT] = t.stat; pr[o + 3 * NT] = gh; pr[o + 4 * NT] = recip_for_pass(gh);
        if (e == steps - 1) {                       // the final sample (no edge starts there): the backward pass starts on it
            const SampleTerms u = prepass_sample<false>(V, A0, A0, w, max_angular_vel, max_angular_accel, kr[D - 1]);
            pr[5 * RS - 3] = u.ak; pr[5 * RS - 2] = u.G; pr[5 * RS - 1] = u.stat;
        }
    }
}

// ---- chunk-speculative forward / backward kernels: one CTA per path, one chunk per thread ------------------------
__device__ __forceinline__ bool same_bits(double a, double b)
{
    return __double_as_longlong(a) == __double_as_longlong(b);
}

// (w_i^2 - w_{i-1}^2) / (2|dtheta|) with IEEE results for every special case, but without sending the warp through the
// slow path of the inlined division when the numerator is 0 (cruise), or the denominator is 0 (two samples snapped
// to the same table entry) or NaN (the "straight" marker of the pre-pass).  Generic path (rare inputs only).
__device__ __noinline__ double accel_ang_div(double num, double h2)
{
    bool special = !(h2 > 0.0) || num == 0.0 || !(fabs(num) < 1e300);
    double n = special ? 1.0 : num, d = special ? 1.0 : h2;
    asm volatile("" : "+d"(n), "+d"(d));   // keep nvcc from folding the selects back into the division's operands
    double q = n / d;
    if (!special) return q;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (h2 > 0.0) {
        if (num == 0.0) return num;                      // +-0 / positive
        return (num != num) ? qnan : copysign(inf, num); // |num| >= 1e300 or inf: overflow / inf / NaN, as IEEE '/' gives
    }
    if (h2 == 0.0) {                                     // x / +0
        if (num > 0.0) return inf;
        if (num < 0.0) return -inf;
        return qnan;                                     // 0/0 or NaN/0
    }
    return qnan;                                         // NaN denominator (the pre-pass's "straight" marker)
}

// Reciprocal of the division's denominator g = 2|dtheta|, made ONCE per sample by the pre-pass (it does not depend on the
// velocity state), so that the state-dependent chain of a pass step holds one multiplication and two fused residual
// corrections instead of a reciprocal refinement:
//    > 0 and finite : RN(1/g), g a normal number whose significand is not all ones (Markstein's exception)
//    +inf           : g == 0 (two samples snapped to the same heading entry; ~44 % of the samples)
//    -1             : anything else (g NaN / subnormal / all-ones significand): the pass takes the generic division
__device__ __forceinline__ double recip_for_pass(double g)
{
    const long long bits = __double_as_longlong(g);
    const bool safe = (g >= 2.2250738585072014e-308) && (g < 1e300) &&
                      ((bits & 0x000FFFFFFFFFFFFFLL) != 0x000FFFFFFFFFFFFFLL);
    double d = safe ? g : 1.0;
    asm volatile("" : "+d"(d));            // the division must never see the special values (warp-wide slow path)
    const double r = 1.0 / d;
    return safe ? r : ((g == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : -1.0);
}

// num / h2 given rce = recip_for_pass(h2) (or NaN when h2 is the straight marker).  q0 = num * r, then two residual
// corrections: after the first the quotient is faithful, and with the correctly rounded reciprocal one more gives the
// correctly rounded quotient (Markstein), i.e. exactly what '/' returns (the sign of a zero quotient is not kept; it
// cannot reach any output).  h2 == 0 and the straight marker fall out of the multiplication: x * inf and x * NaN are
// what x / 0 and x / NaN give.  tests/test_gpu_parity.py::test_recip_division_is_ieee checks 2^30 random pairs.
__device__ __forceinline__ double accel_ang_fast(double num, double h2, double rce)
{
    const double q0 = num * rce;
    double q = fma(fma(-h2, q0, num), rce, q0);
    q = fma(fma(-h2, q, num), rce, q);
    const bool corr = rce <= 1.7976931348623157e308;          // false for NaN and +inf
    const double an = fabs(num);
    const bool in_range = (an < 1e200) && ((an > 1e-200) || (num == 0.0));
    if (rce < 0.0 || (corr && !in_range)) return accel_ang_div(num, h2);     // never taken on sane inputs
    return corr ? q : q0;
}

// One pass step from the hoisted terms.  ak = |kappa| of the step's sample, gh / rg the wheel-acceleration denominator and
// its reciprocal, stat the state-independent acceleration limit, cap the state-independent velocity limit (already merged
// with the initial / forward velocity), sq carries (v|kappa|)^2 of the previous step.
// forward step i -> i+1 (motion_profile_generator.py:193-249): |accel_ang| into the wheel term
__device__ __forceinline__ double fwd_step(double ak, double gh, double rg, double stat, double cap, double v, double& sq,
                                           double acc, double hw, double dd)
{
    const double ang_vel = v * ak;
    const double sqn = ang_vel * ang_vel;
    const double rce = (ak < 1e-6) ? __longlong_as_double(0x7ff8000000000000LL) : rg;   // straight: the term is ignored
    const double accel_ang = accel_ang_fast(sqn - sq, gh, rce);
    const double x = fabs(accel_ang) * hw;                   // ang * w / 2: halving is exact, so (ang*w)/2 == ang*(w/2)
    const double l = acc + x, rr = acc - x;                  // wheel_accel (:52-59)
    double a_wheel = (fabs(l) < fabs(rr)) ? l : rr;
    if (a_wheel < 0) a_wheel = 0;
    const double a = pymin(stat, a_wheel);
    const double s = sqrt(v * v + 2 * a * dd);
    sq = sqn;
    return pymin(cap, s);
}
// backward step i -> i-1 (:255-311): signed accel_ang; cap = min(vel_f[i-1], G_i)
__device__ __forceinline__ double bwd_step(double ak, double gh, double rg, double stat, double cap, double v, double& sq,
                                           double acc, double hw, double dd)
{
    const double ang_vel = v * ak;
    const double sqn = ang_vel * ang_vel;
    const double rce = (ak < 1e-6) ? __longlong_as_double(0x7ff8000000000000LL) : rg;
    const double accel_ang = accel_ang_fast(sqn - sq, gh, rce);
    const double x = accel_ang * hw;
    const double l = acc + x, rr = acc - x;
    double a_wheel = (fabs(l) < fabs(rr)) ? l : rr;
    if (a_wheel < 0) a_wheel = 0;
    const double dcl = pymin(stat, a_wheel);
    const double pv = sqrt(v * v + 2 * dcl * dd);
    sq = sqn;
    return pymin(pv, cap);
}

// ------------------------------------------------------------------------------------------------------------------
// Chunk-speculative passes: CTA = one path, thread c = chunk c (blockDim.x chunks).
//
// Sweep 1: every thread runs its own chunk from a guessed state (lockstep, all lanes busy).  Fix-up rounds: a chunk whose
// predecessor's end state differs bitwise from the state it started from re-runs from the true state; the re-run stops as
// soon as two consecutive velocities equal the stored ones bitwise (from there the old values ARE the serial values).
// Rounds end when no chunk had to re-run; then, by induction from chunk 0, every value is the serial value.
//
// The kernels are bound by the latency of the dependent fp64 chain of one step, so they are written for residency
// (72 registers: 28 one-warp CTAs per SM, 4144 >= 4096 paths in one wave) and a short chain: event tables in shared
// memory, 32-bit indices, records fetched one step ahead into two named buffers (no register rotation), the division's
// reciprocal from the pre-pass, coalesced slot-order streams.
// ------------------------------------------------------------------------------------------------------------------
#define CH_INT_MAX 2147483647

// per-path event tables of the forward pass in shared memory
struct FwdTables {
    const int* bi; const double* acc; int n_b;       // max_acc regimes: acc[j] from sample bi[j] on
    const int* vi; const double* vv; int n_v;        // initial velocity: vv[j] from sample vi[j] on (stops / end included)
};

// forward chunk: edges lo .. lo+len-1; p points at (row 0, field 0, own column) of the record rows, q at row 0 of the forward
// velocities.  Edge e = lo + r reads row r and writes the forward velocity of sample e+1 into row r+1 (the bottom of the chunk
// writes *last: row 0 of the next column, or the tail slot).  NT is a compile-time constant: every access is pointer +
// immediate.
#define RUN_SWEEP 0     // first sweep: plain stores
#define RUN_RERUN 1     // fix-up: bitwise merge detection against the stored velocities
#define RUN_DRY 2       // warm-up before a speculative chunk: state only, no stores
template <int NT, int MODE>
__device__ __forceinline__ bool fwd_run(const double* __restrict__ p, double* __restrict__ q, double* __restrict__ last,
                                        int lo, int len, const FwdTables& T, double hw,
                                        double dd, double& v, double& sq, bool prev_same)
{
    int j = 0;
    while (j + 1 < T.n_b && T.bi[j + 1] <= lo) j++;
    double acc = T.acc[j];
    int nb_next = (j + 1 < T.n_b) ? T.bi[j + 1] : CH_INT_MAX;
    int jv = 0;                                      // initial velocity of sample e+1
    while (jv + 1 < T.n_v && T.vi[jv + 1] <= lo + 1) jv++;
    double v0n = T.vv[jv];
    int nv_next = (jv + 1 < T.n_v) ? T.vi[jv + 1] : CH_INT_MAX;
    double aka = __ldg(p), Ga = __ldg(p + NT), sta = __ldg(p + 2 * NT), gha = __ldg(p + 3 * NT), rga = __ldg(p + 4 * NT);
    double akb, Gb, stb, ghb, rgb;
    constexpr bool RERUN = (MODE == RUN_RERUN);
    constexpr bool DRY = (MODE == RUN_DRY);
    double olda = 0.0, oldb = 0.0;
    if (RERUN) olda = (len > 1) ? q[NT] : *last;
    int e = lo;
    const int hi = lo + len;
#define FWD_ONE(AK_, G_, ST_, GH_, RG_, OLD_)                                                                        \
        if (e == nb_next) { j++; acc = T.acc[j]; nb_next = (j + 1 < T.n_b) ? T.bi[j + 1] : CH_INT_MAX; }              \
        if (e + 1 == nv_next) { jv++; v0n = T.vv[jv]; nv_next = (jv + 1 < T.n_v) ? T.vi[jv + 1] : CH_INT_MAX; }      \
        {                                                                                                            \
            v = fwd_step(AK_, GH_, RG_, ST_, pymin(v0n, G_), v, sq, acc, hw, dd);                                    \
        }                                                                                                            \
        if (RERUN) {                                                                                                 \
            const bool same = same_bits(OLD_, v);                                                                    \
            if (same && prev_same) return true;       /* state equals the old run's: the rest is unchanged */        \
            prev_same = same;                                                                                        \
        }                                                                                                            \
        if (++e >= hi) { if (!DRY) *last = v; break; }                                                               \
        p += 5 * NT;                                                                                                 \
        if (!DRY) { q += NT; *q = v; }
    while (true) {
        // ---- buffers a (look-ahead loads never leave the path's rows: they are padded)
        akb = __ldg(p + 5 * NT); Gb = __ldg(p + 6 * NT); stb = __ldg(p + 7 * NT); ghb = __ldg(p + 8 * NT); rgb = __ldg(p + 9 * NT);
        if (RERUN) oldb = (e + 2 < hi) ? q[2 * NT] : *last;
        FWD_ONE(aka, Ga, sta, gha, rga, olda)
        // ---- buffers b
        aka = __ldg(p + 5 * NT); Ga = __ldg(p + 6 * NT); sta = __ldg(p + 7 * NT); gha = __ldg(p + 8 * NT); rga = __ldg(p + 9 * NT);
        if (RERUN) olda = (e + 2 < hi) ? q[2 * NT] : *last;
        FWD_ONE(akb, Gb, stb, ghb, rgb, oldb)
    }
#undef FWD_ONE
    return false;
}

// Forward pass.  Chunk c = thread c.  vfT: forward velocities in slot order (vfT[slot(e)] = velocity at sample e).
template <int NT>
__global__ void __maxnreg__(72) k_fwd_chunked(
    const int* __restrict__ status, const double* __restrict__ cons, double dd, double start_vel, double end_vel,
    long long RS, const int* __restrict__ n_samples, const double* __restrict__ rec, int E_cap,
    const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, const int* __restrict__ vr_idx, const double* __restrict__ vr_val,
    const int* __restrict__ st_idx, const int* __restrict__ n_vr, double* __restrict__ vfT, int* __restrict__ rounds_out,
    int warm, int max_rounds)
{
    extern __shared__ __align__(16) unsigned char s_mem[];
    const int c = threadIdx.x;
    const long long b = blockIdx.x;
    const int VC = 3 * E_cap + 2;                                  // capacity of the initial-velocity table
    double* s_endv = reinterpret_cast<double*>(s_mem);            // [NT] end state of every chunk: v and (v|k|)^2
    double* s_endw = s_endv + NT;
    double* s_usev = s_endw + NT;                                 // [NT] start state every chunk last used
    double* s_usew = s_usev + NT;
    double* s_acc = s_usew + NT;                                  // [E_cap] max_acc of regime j
    double* s_vv = s_acc + E_cap;                                 // [VC]
    int* s_bi = reinterpret_cast<int*>(s_vv + VC);                // [E_cap] first sample of regime j
    int* s_vi = s_bi + E_cap;                                     // [VC]
    __shared__ int s_nv;
    if (status[b] != ST_OK) return;                               // uniform over the CTA
    const int D = n_samples[b];
    const int steps = D - 1;
    double* vf = vfT + (size_t)b * RS;
    if (c == 0) { vf[0] = start_vel; vf[RS - 1] = start_vel; if (rounds_out) rounds_out[2 * b] = 0; }
    if (steps <= 0) return;
    const int n_b = n_ev[2 * b + 1];
    for (int k = c; k < n_b; k += NT) {
        s_bi[k] = bidx[(size_t)b * E_cap + k];
        s_acc[k] = max_accels[(size_t)b * E_cap + bval[(size_t)b * E_cap + k]];
    }
    if (c == 0) {
        // initial velocity v0[x] as one sorted breakpoint table (motion_profile_generator.py:95-176): the regime value from
        // vr_idx[j] on, 0.01 at the stop samples, end_vel at the last sample.  Three sorted streams (regime starts, stops,
        // sample after a stop) are merged; the value at a breakpoint follows the reference's order of overwrites.
        const int* vi = vr_idx + (size_t)b * E_cap;
        const double* vv = vr_val + (size_t)b * E_cap;
        const int* si = st_idx + (size_t)b * E_cap;
        const int nvr = n_vr[2 * b], nst = n_vr[2 * b + 1];
        int pv = 0, ps = 0, pa = 0, n = 0, reg = 0;
        while (true) {
            int x = CH_INT_MAX;
            if (pv < nvr) x = vi[pv];
            if (ps < nst && si[ps] < x) x = si[ps];
            if (pa < nst && si[pa] + 1 < x) x = si[pa] + 1;
            if (x >= D - 1) break;
            bool stop = false;
            while (pv < nvr && vi[pv] == x) { reg = pv; pv++; }
            while (ps < nst && si[ps] == x) { stop = true; ps++; }
            while (pa < nst && si[pa] + 1 == x) pa++;
            s_vi[n] = x; s_vv[n] = stop ? 0.01 : vv[reg]; n++;
        }
        s_vi[n] = D - 1; s_vv[n] = end_vel; n++;
        s_nv = n;
    }
    const int Lc = chunk_len(steps, NT);
    const int nch = (steps + Lc - 1) / Lc;
    const bool active = c < nch;
    const int lo = c * Lc;
    const int len = (lo + Lc < steps) ? Lc : steps - lo;
    const double* P = rec + (size_t)b * RS * 5;
    double* last = vf + ((lo + len < steps) ? c + 1 : (int)(RS - 1));       // where the velocity of sample lo+len goes
    const double hw = cons[b * 6 + 5] * 0.5;
    __syncthreads();
    FwdTables T;
    T.bi = s_bi; T.acc = s_acc; T.n_b = n_b; T.vi = s_vi; T.vv = s_vv; T.n_v = s_nv;

    // ---- sweep 1: chunk c > 0 starts `wl` steps BEFORE its own range (in the previous column) from the guess "the state-
    // independent caps bind on the two samples before that point": v[s0] = C[s0-1] = min(v0[s0], G[s0-1]) and
    // omega_prev = C[s0-2] |kappa[s0-1]|, and only computes (no stores) until it reaches its range: a wrong guess survives for
    // at most one acceleration ramp, so after the warm-up the state is usually the true one and the fix-up rounds have little
    // to re-run.  (Exactness never depends on the guess: the rounds below verify bitwise.)
    {
        double v = start_vel, sq = 0.0;
        if (active) {
            if (c > 0) {
                auto v0_at = [&](int x) { int q = 0; while (q + 1 < T.n_v && T.vi[q + 1] <= x) q++; return T.vv[q]; };
                auto term = [&](int x, int fld) { const int cx = x / Lc, rx = x - cx * Lc; return __ldg(P + ((size_t)rx * 5 + fld) * NT + cx); };
                int wl = warm < Lc ? warm : Lc;                      // the warm-up stays inside the previous column
                if (wl > lo - 2) wl = lo - 2 > 0 ? lo - 2 : 0;
                const int s0 = lo - wl;
                const double vm1 = (s0 >= 2) ? pymin(v0_at(s0 - 1), term(s0 - 2, 1)) : start_vel;
                v = pymin(v0_at(s0), term(s0 - 1, 1));
                const double wp = vm1 * term(s0 - 1, 0);
                sq = wp * wp;
                if (wl > 0) fwd_run<NT, RUN_DRY>(P + (size_t)(Lc - wl) * 5 * NT + c - 1, nullptr, nullptr, s0, wl, T, hw, dd, v, sq, false);
            }
            s_usev[c] = v; s_usew[c] = sq;
            fwd_run<NT, RUN_SWEEP>(P + c, vf + c, last, lo, len, T, hw, dd, v, sq, false);
        }
        s_endv[c] = v; s_endw[c] = sq;
    }
    __syncthreads();

    // ---- fix-up rounds
    int rounds = 0;
    for (int round = 1; round < NT && round <= max_rounds; round++) {
        bool need = false;
        double in_v = 0.0, in_w = 0.0;
        if (active && c >= round) {
            in_v = s_endv[c - 1]; in_w = s_endw[c - 1];
            need = !(same_bits(in_v, s_usev[c]) && same_bits(in_w, s_usew[c]));
        }
        if (!__syncthreads_or(need)) break;          // also orders this round's reads before its writes
        rounds = round;
        if (need) {
            double v = in_v, sq = in_w;
            const bool merged = fwd_run<NT, RUN_RERUN>(P + c, vf + c, last, lo, len, T, hw, dd, v, sq,
                                                  same_bits(in_v, s_usev[c]));
            s_usev[c] = in_v; s_usew[c] = in_w;
            if (!merged) { s_endv[c] = v; s_endw[c] = sq; }
        }
        __syncthreads();
    }
    if (c == 0 && rounds_out) rounds_out[2 * b] = rounds;
}

// backward chunk: edges lo+len-1 down to lo.  p points at (row len-1, field 0, own column) of the record rows, f / o at the
// same row of the forward / final velocities.  Edge e = lo + r (the reference's step i = e+1 -> e) uses the terms of sample
// e+1 (fields 0-2 of row r+1; for the chunk's top edge the three `top` values), gh / rg (fields 3-4) and the forward velocity
// of row r, and writes the final velocity of sample e into row r of the final velocities.
// OVR: the path has max_acceleration overrides; the static limit of the backward pass then comes from its own slot-order
// array `sb` (same slots as the forward velocities) instead of field 2 of the record rows.
template <int NT, int MODE, bool OVR>
__device__ __forceinline__ bool bwd_run(const double* __restrict__ p, const double* __restrict__ f, double* __restrict__ o,
                                        const double* __restrict__ sb, int lo, int len, double top_ak, double top_G,
                                        double top_st, const int* s_bi, const double* s_acc, int n_b, double acc0, double hw,
                                        double dd, double& v, double& sq, bool prev_same)
{
    // regime at the chunk start (walking down from D-1): the smallest boundary index > i was the last one applied
    int e = lo + len - 1;                            // current edge; the reference's loop index is i = e + 1
    int j = n_b - 1;
    double acc = acc0;
    while (j >= 0 && s_bi[j] > e + 1) { acc = s_acc[j]; j--; }
    int nb_next = (j >= 0) ? s_bi[j] : -1;
    double aka = top_ak, Ga = top_G, sta = top_st, akb, Gb, stb;     // terms of sample e+1
    double gha = __ldg(p + 3 * NT), rga = __ldg(p + 4 * NT), ghb, rgb;
    constexpr bool RERUN = (MODE == RUN_RERUN);
    constexpr bool DRY = (MODE == RUN_DRY);
    double fa = __ldg(f), fb;
    double olda = 0.0, oldb = 0.0;
    if (RERUN) olda = *o;
#define BWD_ONE(AK_, G_, ST_, GH_, RG_, F_, OLD_)                                                                    \
        if (e + 1 == nb_next) { acc = s_acc[j]; j--; nb_next = (j >= 0) ? s_bi[j] : -1; }                             \
        v = bwd_step(AK_, GH_, RG_, ST_, pymin(F_, G_), v, sq, acc, hw, dd);                                         \
        if (RERUN) {                                                                                                 \
            const bool same = same_bits(OLD_, v);                                                                    \
            if (same && prev_same) return true;                                                                      \
            prev_same = same;                                                                                        \
        }                                                                                                            \
        if (!DRY) *o = v;                                                                                            \
        if (--e < lo) break;                                                                                         \
        p -= 5 * NT; f -= NT;                                                                                        \
        if (OVR) sb -= NT;                                                                                           \
        if (!DRY) o -= NT;
    while (true) {
        // ---- buffers a: the next step (edge e-1) needs the terms of sample e (this row) and gh / rg / vf / old of the row
        // below (at the chunk's first row the look-ahead stays on the row: the values are not used)
        akb = __ldg(p); Gb = __ldg(p + NT); stb = OVR ? __ldg(sb) : __ldg(p + 2 * NT);
        if (e > lo) { ghb = __ldg(p - 2 * NT); rgb = __ldg(p - NT); fb = __ldg(f - NT); if (RERUN) oldb = o[-NT]; }
        else { ghb = 0.0; rgb = 0.0; fb = 0.0; }
        BWD_ONE(aka, Ga, sta, gha, rga, fa, olda)
        // ---- buffers b
        aka = __ldg(p); Ga = __ldg(p + NT); sta = OVR ? __ldg(sb) : __ldg(p + 2 * NT);
        if (e > lo) { gha = __ldg(p - 2 * NT); rga = __ldg(p - NT); fa = __ldg(f - NT); if (RERUN) olda = o[-NT]; }
        else { gha = 0.0; rga = 0.0; fa = 0.0; }
        BWD_ONE(akb, Gb, stb, ghb, rgb, fb, oldb)
    }
#undef BWD_ONE
    return false;
}

// Backward pass.  Thread k walks column nch-1-k (the k-th chunk counted from the end), so states flow from thread k-1 to
// thread k as in the forward kernel.  Reads the forward velocities and writes the final ones, both in slot order
// (velT[RS-1] = end_vel is the last sample); k_untranspose puts them into sample order.  Also accumulates the
// travel-time estimate used to size the time-domain outputs.
template <int NT>
__global__ void __maxnreg__(72) k_bwd_chunked(
    const int* __restrict__ status, const double* __restrict__ cons, double dd, double dt, double end_vel,
    long long RS, const int* __restrict__ n_samples, const double* __restrict__ rec, int E_cap,
    const double* __restrict__ max_accels, const int* __restrict__ bidx, const int* __restrict__ bval,
    const int* __restrict__ n_ev, const double* __restrict__ vfT, double* __restrict__ velT, float* __restrict__ t_est,
    int* __restrict__ rounds_out, int warm, int max_rounds, const double* __restrict__ statB)
{
    extern __shared__ __align__(16) unsigned char s_mem[];
    const int k = threadIdx.x;
    const long long b = blockIdx.x;
    double* s_endv = reinterpret_cast<double*>(s_mem);
    double* s_endw = s_endv + NT;
    double* s_usev = s_endw + NT;
    double* s_usew = s_usev + NT;
    double* s_acc = s_usew + NT;                                  // [E_cap] max_accels[bval[j] + 1] (applied at sample bidx[j])
    int* s_bi = reinterpret_cast<int*>(s_acc + E_cap);
    if (k == 0 && t_est) t_est[b] = 0.f;
    if (status[b] != ST_OK) return;
    const int D = n_samples[b];
    const int steps = D - 1;
    double* vo = velT + (size_t)b * RS;
    if (k == 0) { vo[RS - 1] = end_vel; if (rounds_out) rounds_out[2 * b + 1] = 0; }
    if (steps <= 0) return;
    const int n_b = n_ev[2 * b + 1];
    for (int q = k; q < n_b; q += NT) {
        s_bi[q] = bidx[(size_t)b * E_cap + q];
        s_acc[q] = max_accels[(size_t)b * E_cap + bval[(size_t)b * E_cap + q] + 1];
    }
    const double acc0 = max_accels[(size_t)b * E_cap + bval[(size_t)b * E_cap + n_b - 1]];
    const int Lc = chunk_len(steps, NT);
    const int nch = (steps + Lc - 1) / Lc;
    const bool active = k < nch;
    const int col = nch - 1 - k;                                   // column (forward chunk index)
    const int lo = col * Lc;
    const int len = (lo + Lc < steps) ? Lc : steps - lo;
    const double* P = rec + (size_t)b * RS * 5;
    const double* vf = vfT + (size_t)b * RS;
    const double hw = cons[b * 6 + 5] * 0.5;
    // with max_acceleration overrides the static limit of this pass is in statB (slot order, written by the pre-pass)
    const bool ovr = pass_has_override(max_accels + (size_t)b * E_cap, n_ev[2 * b], cons[b * 6 + 1]) != 0;
    const double* SB = statB + (size_t)b * RS;
    __syncthreads();

    // terms of the sample above the chunk (sample lo+len): row 0 of the next column, or the tail
    double top_ak = 0.0, top_G = 0.0, top_st = 0.0;
    if (active) {
        if (lo + len < steps) {
            top_ak = __ldg(P + col + 1); top_G = __ldg(P + NT + col + 1);
            top_st = ovr ? __ldg(SB + col + 1) : __ldg(P + 2 * NT + col + 1);
        } else {
            top_ak = __ldg(P + 5 * RS - 3); top_G = __ldg(P + 5 * RS - 2);
            top_st = ovr ? __ldg(SB + RS - 1) : __ldg(P + 5 * RS - 1);
        }
    }
    const size_t r_top = active ? (size_t)(len - 1) : 0;           // row of the chunk's top edge
    const double* p0 = P + r_top * 5 * NT + (active ? col : 0);
    const double* f0 = vf + r_top * NT + (active ? col : 0);
    const double* sb0 = SB + r_top * NT + (active ? col : 0);
    double* o0 = vo + r_top * NT + (active ? col : 0);

    // ---- sweep 1: thread k > 0 starts `wl` steps ABOVE its chunk (in the next column) from the guess
    // v[s0] = min(vel_f[s0], G[s0+1]) (the state-independent part of what step s0+1 produces) and only computes until it
    // reaches its own range (see the forward kernel)
    {
        double v = end_vel, sq = 0.0;
        if (active) {
            if (k > 0) {
                const int hi = lo + len;                                   // sample at the top of this chunk (< D-1)
                auto rec_of = [&](int x, int fld) {                         // field fld of sample x
                    if (fld == 2 && ovr) return __ldg(SB + (x >= steps ? (int)(RS - 1) : edge_slot(x, Lc, NT)));
                    if (x >= steps) return __ldg(P + 5 * RS - 3 + fld);
                    const int cx = x / Lc, rx = x - cx * Lc;
                    return __ldg(P + ((size_t)rx * 5 + fld) * NT + cx);
                };
                const int len_up = (hi + Lc < steps) ? Lc : steps - hi;    // edges of the column above
                int wl = warm < len_up ? warm : len_up;
                if (hi + wl + 2 > D - 1) wl = (D - 1) - hi - 2 > 0 ? (D - 1) - hi - 2 : 0;
                const int s0 = hi + wl;                                    // sample the guess is made at
                const int s1 = edge_slot(s0, Lc, NT);
                v = pymin(vf[s1], rec_of(s0 + 1, 1));
                double vp1 = end_vel;
                if (s0 + 2 <= D - 1) vp1 = pymin(vf[edge_slot(s0 + 1, Lc, NT)], rec_of(s0 + 2, 1));
                const double wp = vp1 * rec_of(s0 + 1, 0);
                sq = wp * wp;
                if (wl > 0) {
                    // edges s0-1 .. hi of column col+1 (rows wl-1 .. 0); the terms above the first of them are sample s0's
                    const size_t rw = (size_t)(wl - 1);
                    const double* pd = P + rw * 5 * NT + col + 1;
                    const double* fd = vf + rw * NT + col + 1;
                    const double* sd = SB + rw * NT + col + 1;
                    if (ovr) bwd_run<NT, RUN_DRY, true>(pd, fd, nullptr, sd, hi, wl, rec_of(s0, 0), rec_of(s0, 1), rec_of(s0, 2),
                                                        s_bi, s_acc, n_b, acc0, hw, dd, v, sq, false);
                    else bwd_run<NT, RUN_DRY, false>(pd, fd, nullptr, sd, hi, wl, rec_of(s0, 0), rec_of(s0, 1), rec_of(s0, 2),
                                                     s_bi, s_acc, n_b, acc0, hw, dd, v, sq, false);
                }
            }
            s_usev[k] = v; s_usew[k] = sq;
            if (ovr) bwd_run<NT, RUN_SWEEP, true>(p0, f0, o0, sb0, lo, len, top_ak, top_G, top_st, s_bi, s_acc, n_b, acc0, hw, dd,
                                                  v, sq, false);
            else bwd_run<NT, RUN_SWEEP, false>(p0, f0, o0, sb0, lo, len, top_ak, top_G, top_st, s_bi, s_acc, n_b, acc0, hw, dd,
                                               v, sq, false);
        }
        s_endv[k] = v; s_endw[k] = sq;
    }
    __syncthreads();

    // ---- fix-up rounds (states flow from thread k-1 to thread k, as in the forward kernel)
    int rounds = 0;
    for (int round = 1; round < NT && round <= max_rounds; round++) {
        bool need = false;
        double in_v = 0.0, in_w = 0.0;
        if (active && k >= round) {
            in_v = s_endv[k - 1]; in_w = s_endw[k - 1];
            need = !(same_bits(in_v, s_usev[k]) && same_bits(in_w, s_usew[k]));
        }
        if (!__syncthreads_or(need)) break;
        rounds = round;
        if (need) {
            double v = in_v, sq = in_w;
            const bool ps = same_bits(in_v, s_usev[k]);
            const bool merged = ovr ? bwd_run<NT, RUN_RERUN, true>(p0, f0, o0, sb0, lo, len, top_ak, top_G, top_st, s_bi, s_acc,
                                                                   n_b, acc0, hw, dd, v, sq, ps)
                                    : bwd_run<NT, RUN_RERUN, false>(p0, f0, o0, sb0, lo, len, top_ak, top_G, top_st, s_bi, s_acc,
                                                                    n_b, acc0, hw, dd, v, sq, ps);
            s_usev[k] = in_v; s_usew[k] = in_w;
            if (!merged) { s_endv[k] = v; s_endw[k] = sq; }
        }
        __syncthreads();
    }
    if (k == 0 && rounds_out) rounds_out[2 * b + 1] = rounds;

    // ---- travel-time estimate (single precision is plenty: it only sizes buffers)
    __syncthreads();
    float est = 0.f;
    if (active) {
        const int top = (lo + len < steps) ? col + 1 : (int)(RS - 1);       // slot of the sample above the chunk
        float vprev = (float)vo[top];
        for (int sl = (len - 1) * NT + col; sl >= 0; sl -= NT) {
            const float vcur = (float)vo[sl];
            const float vm = 0.5f * (vprev + vcur);
            est += __fdividef((float)dd, fmaxf(vm, 0.05f) * (float)dt);
            vprev = vcur;
        }
    }
    float* s_f = reinterpret_cast<float*>(s_endv);      // end states are no longer needed
    s_f[k] = est;
    __syncthreads();
    if (k == 0 && t_est) {
        float tot = 0.f;
        for (int q = 0; q < NT; q++) tot += s_f[q];
        t_est[b] = tot;
    }
}

// slot order -> sample order through a 32 x 32 shared-memory tile (both sides coalesced): vel[e] = vT[slot(e)], and the
// last sample from slot RS-1.  grid = (row tiles * column tiles, B), 256 threads.
__global__ void __launch_bounds__(256) k_untranspose(const int* __restrict__ status, const int* __restrict__ n_samples,
                                                     long long D_cap, long long RS, int NT,
                                                     const double* __restrict__ vT, double* __restrict__ vel,
                                                     unsigned tiles_x)
{
    __shared__ double tile[32][33];
    const PathTile pt = path_tile(tiles_x);
    const long long b = pt.b;
    if (status[b] != ST_OK) return;
    const int D = n_samples[b];
    const int steps = D - 1;
    const double* src = vT + (size_t)b * RS;
    double* dst = vel + (size_t)b * D_cap;
    if (pt.x == 0 && threadIdx.x == 0) dst[D - 1] = src[RS - 1];
    if (steps <= 0) return;
    const int Lc = chunk_len(steps, NT);
    const int ctiles = (NT + 31) >> 5;
    const int rt = pt.x / ctiles, ct = pt.x - rt * ctiles;
    const int s0 = rt * 32, c0 = ct * 32;
    if (s0 >= Lc) return;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
    for (int r = wrp; r < 32; r += 8) {
        const int s = s0 + r;
        tile[r][lane] = (s < Lc && c0 + lane < NT) ? src[(size_t)s * NT + c0 + lane] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = wrp; q < 32; q += 8) {
        const int c = c0 + q;
        const int s = s0 + lane;
        const int e = c * Lc + s;
        if (s < Lc && c < NT && e < steps) dst[e] = tile[lane][q];
    }
}
