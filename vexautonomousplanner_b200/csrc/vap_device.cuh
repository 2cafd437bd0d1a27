// vap_device.cuh -- device-side arithmetic shared by every kernel of libvap.so (sm_100a).
//
// Everything here must reproduce the reference's IEEE-754 binary64 arithmetic operation by operation:
// the translation unit is compiled with -fmad=false, so '*' followed by '+' stays two roundings, and
// fma() is written explicitly only where the reference's BLAS fuses (1-D np.linalg.norm, 2x2 matmul).
// Citations are file:line relative to the reference's src/.
#pragma once
#include <cstdint>
#include <math.h>

#define NA 12
enum { A_X = 0, A_Y, A_TURN, A_WAIT, A_MAXVEL, A_MAXACC, A_TX, A_TY, A_INMAG, A_OUTMAG, A_RCOS, A_RSIN };
#define F_REVERSE 1
#define F_STOP 2
#define F_TANGENT 4
#define APA 4
enum { P_T = 0, P_WAIT, P_MAXVEL, P_MAXACC };

#define ST_OK 0
#define ST_FALSE (-1)
#define ST_INDEX (-2)
#define ST_VALUE (-3)
#define ST_CAPACITY (-4)
#define ST_DIVERGED (-5)     // the reference's loop would not terminate (or would emit more than VAP_ROW_LIMIT rows)
#define ST_EVENTS (-6)       // more node-crossing / action-point candidates than the event tables hold (never goes away on a retry)
#define VAP_ROW_LIMIT 50000000LL

#define VAP_PI 3.141592653589793

// cp.async (LDGSTS): 16-byte asynchronous global -> shared copies, tracked per thread in commit groups
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier): one instruction moves a whole
// contiguous block (a 1 KB block of a time-loop thread's ring, a 1 KB plane of a pass warp's block row) instead of 64 LDGSTS.
__device__ __forceinline__ void mbar_init(unsigned a, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned a, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}" ::"r"(a), "r"(parity) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy reads of a shared-memory slot must precede the async-proxy (TMA) writes that refill it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- bounds diagnostics (-DVAP_BOUNDS_CHECK; compute-sanitizer is closed on the GPU pool this was developed on) ------------
// Every hand-computed index of the TMA rings, the chunk-interleaved pass arrays and the sampling tiles is tested against the
// extent of its array before the access; a violation is counted per site (vap_diag_read) instead of faulting.  The release
// build compiles the checks away.
#ifdef VAP_BOUNDS_CHECK
#define VAP_DIAG_SITES 32
__device__ unsigned long long g_vap_diag[VAP_DIAG_SITES];
#define VAP_CHECK(site, cond) do { if (!(cond)) atomicAdd(&g_vap_diag[site], 1ull); else atomicAdd(&g_vap_diag[16 + ((site) & 15)], 1ull); } while (0)
#else
#define VAP_CHECK(site, cond) do { } while (0)
#endif

// Kernels whose grid is (tiles of one path) x (paths) are launched on a 1-D grid of tiles_x * B CTAs (grid.y stops at
// 65535 paths); a CTA finds its path and its tile with one division.  Tiles of a path stay adjacent in launch order.
struct PathTile { long long b; unsigned x; };
__device__ __forceinline__ PathTile path_tile(unsigned tiles_x)
{
    PathTile p;
    const unsigned q = blockIdx.x / tiles_x;
    p.b = q; p.x = blockIdx.x - q * tiles_x;
    return p;
}

// Python builtin min / max on floats: the first argument survives unless a later one is strictly smaller / larger.
__device__ __forceinline__ double pymin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double pymax(double a, double b) { return (b > a) ? b : a; }
// np.linalg.norm on a length-2 vector (OpenBLAS fuses the second product) vs. norm(axis=1) (plain).
__device__ __forceinline__ double norm1d(double x, double y) { return sqrt(fma(y, y, x * x)); }
__device__ __forceinline__ double norm_ax1(double x, double y) { return sqrt(x * x + y * y); }
// t % 1 for Python floats / numpy (result in [0,1)); t - floor(t) is the same single rounding.
__device__ __forceinline__ double frac1(double t) { return t - floor(t); }
// Python float % with positive divisor y
__device__ __forceinline__ double pymod_pos(double x, double y)
{
    double m = fmod(x, y);
    if (m != 0.0) { if (m < 0) m += y; } else m = 0.0;
    return m;
}

// ---- quintic Hermite basis, expanded monomials summed left to right (quintic_hermite_spline.py:288-416)
template <int WHICH>
__device__ __forceinline__ void basis(double t, double* H)
{
    double t2 = t * t, t3 = t2 * t;
    if (WHICH == 0) {
        double t4 = t3 * t, t5 = t4 * t;
        H[0] = 1 - 10 * t3 + 15 * t4 - 6 * t5;
        H[1] = 10 * t3 - 15 * t4 + 6 * t5;
        H[2] = t - 6 * t3 + 8 * t4 - 3 * t5;
        H[3] = -4 * t3 + 7 * t4 - 3 * t5;
        H[4] = 0.5 * t2 - 1.5 * t3 + 1.5 * t4 - 0.5 * t5;
        H[5] = 0.5 * t3 - t4 + 0.5 * t5;
    } else if (WHICH == 1) {
        double t4 = t3 * t;
        H[0] = -30 * t2 + 60 * t3 - 30 * t4;
        H[1] = 30 * t2 - 60 * t3 + 30 * t4;
        H[2] = 1 - 18 * t2 + 32 * t3 - 15 * t4;
        H[3] = -12 * t2 + 28 * t3 - 15 * t4;
        H[4] = t - 4.5 * t2 + 6 * t3 - 2.5 * t4;
        H[5] = 1.5 * t2 - 4 * t3 + 2.5 * t4;
    } else {
        H[0] = -60 * t + 180 * t2 - 120 * t3;
        H[1] = 60 * t - 180 * t2 + 120 * t3;
        H[2] = -36 * t + 96 * t2 - 60 * t3;
        H[3] = -24 * t + 84 * t2 - 60 * t3;
        H[4] = 1 - 9 * t + 18 * t2 - 10 * t3;
        H[5] = 3 * t - 12 * t2 + 10 * t3;
    }
}

// _normalize_parameter (quintic_hermite_spline.py:506-541): clamp, segment index, local parameter
__device__ __forceinline__ void normalize_param(double t, int nseg, double pend, int& idx, double& u)
{
    double tt = pymax(0.0, pymin(t, pend));
    idx = (int)tt;
    if (idx == nseg) idx = nseg - 1;
    u = tt - (double)idx;
}

// spline-level evaluation (get_point / get_derivative / get_second_derivative)
template <int WHICH>
__device__ __forceinline__ void eval_spline(const double* __restrict__ seg, int nseg, double pend, double t,
                                            double& ox, double& oy)
{
    int idx; double u;
    normalize_param(t, nseg, pend, idx, u);
    double H[6];
    basis<WHICH>(u, H);
    const double2* s = reinterpret_cast<const double2*>(seg + (size_t)idx * 12);
    double ax = 0.0, ay = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) { double2 r = __ldg(s + i); ax += H[i] * r.x; ay += H[i] * r.y; }
    ox = ax; oy = ay;
}

// first and second derivative at once (same normalisation; used by the property tables)
__device__ __forceinline__ void eval_spline_d12(const double* __restrict__ seg, int nseg, double pend, double t,
                                                double& dx, double& dy, double& ddx, double& ddy)
{
    int idx; double u;
    normalize_param(t, nseg, pend, idx, u);
    double H1[6], H2[6];
    basis<1>(u, H1);
    basis<2>(u, H2);
    const double2* s = reinterpret_cast<const double2*>(seg + (size_t)idx * 12);
    double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double2 r = __ldg(s + i);
        ax += H1[i] * r.x; ay += H1[i] * r.y;
        bx += H2[i] * r.x; by += H2[i] * r.y;
    }
    dx = ax; dy = ay; ddx = bx; ddy = by;
}

// _map_parameter_to_spline (spline_manager.py:243-275): first spline with t <= its last node index
struct PathGeo {
    const double* seg;      // [G][12]
    const int* first;       // [S+1]
    const double* pend;     // [S]
    int S;
};
__device__ __forceinline__ int map_spline(const PathGeo& g, double t)
{
    int k = 0;
    if (g.S <= 8) {
        while (k < g.S - 1 && !(t <= (double)g.first[k + 1])) k++;
    } else {
        int lo = 0, hi = g.S - 1;   // smallest k with t <= first[k+1], else S-1
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (t <= (double)g.first[mid + 1]) hi = mid; else lo = mid + 1;
        }
        k = lo;
    }
    return k;
}
__device__ __forceinline__ PathGeo path_geo(long long b, int N_max, const double* seg, const int* first_node,
                                            const double* param_end, const int* n_splines)
{
    PathGeo g;
    g.seg = seg + (size_t)b * (N_max - 1) * 12;
    g.first = first_node + (size_t)b * (N_max + 1);
    g.pend = param_end + (size_t)b * N_max;
    g.S = n_splines[b];
    return g;
}

template <int WHICH>
__device__ __forceinline__ void eval_path(const PathGeo& g, double t, double& ox, double& oy)
{
    int k = map_spline(g, t);
    int f0 = g.first[k], f1 = g.first[k + 1];
    eval_spline<WHICH>(g.seg + (size_t)f0 * 12, f1 - f0, g.pend[k], t - (double)f0, ox, oy);
}
__device__ __forceinline__ void eval_path_d12(const PathGeo& g, double t, double& dx, double& dy, double& ddx,
                                              double& ddy)
{
    int k = map_spline(g, t);
    int f0 = g.first[k], f1 = g.first[k + 1];
    eval_spline_d12(g.seg + (size_t)f0 * 12, f1 - f0, g.pend[k], t - (double)f0, dx, dy, ddx, ddy);
}

// curvature / heading from derivatives (spline_manager.py:516-536).  speed2**1.5 is evaluated as
// s*sqrt(s) (<= 1 ulp from the true value, like numpy's SIMD pow; value-only, never an index).
__device__ __forceinline__ void curv_heading(double dx, double dy, double ddx, double ddy, double& kap, double& th)
{
    double s2 = dx * dx + dy * dy;
    double num = dx * ddy - dy * ddx;
    kap = (s2 >= 1e-10) ? num / (s2 * sqrt(s2)) : 0.0;
    th = atan2(dy, dx);
}

// ---- tables -----------------------------------------------------------------------------------------
// np.linspace(0, n-1, P)[j]  (spline_manager.py:487)
__device__ __forceinline__ double prop_param(long long j, long long P, int n, double step)
{
    return (j == P - 1) ? (double)(n - 1) : (double)j * step;
}
// np.searchsorted(params, t) on the analytic grid: O(1) estimate + exact fix-up
__device__ __forceinline__ long long prop_lower_bound(double t, long long P, int n, double step, double inv_step)
{
    double e = t * inv_step;
    long long j = (e > 0.0) ? (e < (double)P ? (long long)e : P) : 0;
    while (j > 0 && prop_param(j - 1, P, n, step) >= t) j--;
    while (j < P && prop_param(j, P, n, step) < t) j++;
    return j;
}
// _interpolate_property (spline_manager.py:550-580): returns the table index to gather (snap branch), or
// the interpolation weights when the (never observed) lerp branch applies.
__device__ __forceinline__ double snap_gather(const double* __restrict__ vals, double t, long long P, int n,
                                              double step, double inv_step)
{
    long long idx = prop_lower_bound(t, P, n, step, inv_step);
    if (idx == 0) return vals[0];
    if (idx >= P) return vals[P - 1];
    double t0 = prop_param(idx - 1, P, n, step), t1 = prop_param(idx, P, n, step);
    if (frac1(t0) != frac1(t1)) return (frac1(t) > 0.5) ? vals[idx - 1] : vals[idx];
    double v0 = vals[idx - 1], v1 = vals[idx];
    return v0 + (v1 - v0) * (t - t0) / (t1 - t0);
}
// two tables at once (same index)
__device__ __forceinline__ void snap_gather2(const double* __restrict__ ka, const double* __restrict__ ha, double t,
                                             long long P, int n, double step, double inv_step, double& k, double& h)
{
    long long idx = prop_lower_bound(t, P, n, step, inv_step);
    if (idx == 0) { k = ka[0]; h = ha[0]; return; }
    if (idx >= P) { k = ka[P - 1]; h = ha[P - 1]; return; }
    double t0 = prop_param(idx - 1, P, n, step), t1 = prop_param(idx, P, n, step);
    if (frac1(t0) != frac1(t1)) {
        long long g = (frac1(t) > 0.5) ? idx - 1 : idx;
        k = ka[g]; h = ha[g];
        return;
    }
    double w = (t - t0);
    k = ka[idx - 1] + (ka[idx] - ka[idx - 1]) * w / (t1 - t0);
    h = ha[idx - 1] + (ha[idx] - ha[idx - 1]) * w / (t1 - t0);
}

// distance_to_time (spline_manager.py:291-318); hint = previous index (or <0), updated on return
__device__ __forceinline__ double distance_to_time(const double* __restrict__ ld, const double* __restrict__ lt,
                                                   long long Q, double total, int n, double d, long long& hint)
{
    if (d <= 0) return 0.0;
    if (d >= total) return (double)(n - 1);
    long long idx;
    long long h = hint;
    // fast path: a monotone walk lands in the same or one of the next intervals
    if (h >= 1 && h < Q && ld[h - 1] < d) {
        idx = h;
        int steps = 0;
        while (idx < Q && ld[idx] < d && steps < 4) { idx++; steps++; }
        if (!(idx < Q && ld[idx] >= d)) h = -1;
    } else h = -1;
    if (h < 0) {
        long long lo = 0, hi = Q;   // np.searchsorted side='left'
        while (lo < hi) {
            long long mid = lo + ((hi - lo) >> 1);
            if (ld[mid] < d) lo = mid + 1; else hi = mid;
        }
        idx = lo;
    }
    hint = idx;
    if (idx == 0) return lt[0];
    double d0 = ld[idx - 1], d1 = ld[idx], t0 = lt[idx - 1], t1 = lt[idx];
    return t0 + (t1 - t0) * (d - d0) / (d1 - d0);
}

// ---- 32-bit index variants for the sample-parallel kernels (tables have < 2^31 entries) ---------------------------
struct PropGrid {            // per-path constants of the analytic parameter grid np.linspace(0, n-1, P)
    int P, n;
    double step, inv_step;
};
__device__ __forceinline__ PropGrid prop_grid(int spn, int n)
{
    PropGrid g;
    g.P = spn * n; g.n = n;
    g.step = (double)(n - 1) / (double)(g.P - 1);
    g.inv_step = 1.0 / g.step;
    return g;
}
__device__ __forceinline__ double prop_param32(int j, const PropGrid& g)
{
    return (j == g.P - 1) ? (double)(g.n - 1) : (double)j * g.step;
}
__device__ __forceinline__ void snap_gather2_32(const double* __restrict__ ka, const double* __restrict__ ha, double t,
                                                const PropGrid& g, double& k, double& h)
{
    double e = t * g.inv_step;
    int j = (e > 0.0) ? (e < (double)g.P ? __double2int_rz(e) : g.P) : 0;
    while (j > 0 && prop_param32(j - 1, g) >= t) j--;
    while (j < g.P && prop_param32(j, g) < t) j++;
    if (j == 0) { k = ka[0]; h = ha[0]; return; }
    if (j >= g.P) { k = ka[g.P - 1]; h = ha[g.P - 1]; return; }
    double t0 = prop_param32(j - 1, g), t1 = prop_param32(j, g);
    if (frac1(t0) != frac1(t1)) {
        int q = (frac1(t) > 0.5) ? j - 1 : j;
        k = ka[q]; h = ha[q];
        return;
    }
    double w = (t - t0);
    k = ka[j - 1] + (ka[j] - ka[j - 1]) * w / (t1 - t0);
    h = ha[j - 1] + (ha[j] - ha[j - 1]) * w / (t1 - t0);
}
__device__ __forceinline__ double distance_to_time32(const double* __restrict__ ld, const double* __restrict__ lt, int Q,
                                                     double total, int n, double d)
{
    if (d <= 0) return 0.0;
    if (d >= total) return (double)(n - 1);
    int lo = 0, hi = Q;      // np.searchsorted side='left'
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(ld + mid) < d) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return lt[0];
    double d0 = ld[lo - 1], d1 = ld[lo], t0 = lt[lo - 1], t1 = lt[lo];
    return t0 + (t1 - t0) * (d - d0) / (d1 - d0);
}

// ---- inverse index of the distance table ("which LUT interval does distance d fall in", in O(1)) ----------------------
// Row layout (int32): [0..1] the bits of scale = Q / total_length, [2 + m] = searchsorted(distances, m / scale) for
// m = 0 .. Q (the lower edge of bucket m).  The index only seeds the search: the result is verified against the
// table in both directions, so it is exactly np.searchsorted(distances, d) (side='left') whatever the seed.
#define LUT_INV_HDR 2
__device__ __forceinline__ double distance_to_time_inv(const double* __restrict__ ld, const double* __restrict__ lt,
                                                       const int* __restrict__ inv, int Q, double total, int n, double d)
{
    if (d <= 0) return 0.0;
    if (d >= total) return (double)(n - 1);
    const double scale = __hiloint2double(__ldg(inv + 1), __ldg(inv));
    const double e = d * scale;
    int m = (e < (double)Q) ? __double2int_rz(e) : Q;
    int lo = __ldg(inv + LUT_INV_HDR + m);
    double d1 = __ldg(ld + lo);
    while (lo < Q - 1 && d1 < d) { lo++; d1 = __ldg(ld + lo); }          // d < total = distances[Q-1]: stops inside the table
    if (lo == 0) return lt[0];
    double d0 = __ldg(ld + lo - 1);
    while (lo > 1 && d0 >= d) { lo--; d1 = d0; d0 = __ldg(ld + lo - 1); }
    if (d0 >= d) return lt[0];                                            // lo == 1 and distances[0] >= d (cannot happen for d > 0)
    const double t0 = __ldg(lt + lo - 1), t1 = __ldg(lt + lo);
    return t0 + (t1 - t0) * (d - d0) / (d1 - d0);
}

// lerp on xs[i] = fl(i*dd) (motion_profile_generator.py:349-386,:484): index of searchsorted(side='right') - 1
__device__ __forceinline__ long long uniform_index(double x, double dd, double inv_dd, long long D)
{
    double e = x * inv_dd;
    long long k = (e > 0.0) ? (e < (double)(D - 1) ? (long long)e : D - 1) : 0;
    while (k + 1 < D && (double)(k + 1) * dd <= x) k++;
    while (k >= 0 && (double)k * dd > x) k--;
    return k;
}
__device__ __forceinline__ double lerp_uniform(double x, double dd, double inv_dd, long long D,
                                               const double* __restrict__ ys)
{
    long long idx = uniform_index(x, dd, inv_dd, D);
    if (idx < 0) return ys[0];
    if (idx >= D - 1) return ys[D - 1];
    double x0 = (double)idx * dd, x1 = (double)(idx + 1) * dd;
    double y0 = ys[idx], y1 = ys[idx + 1];
    return y0 + (x - x0) * (y1 - y0) / (x1 - x0);
}

// ---- velocity-pass pieces (motion_profile_generator.py:23-59) ----------------------------------------
__device__ __forceinline__ double wheel_accel(double acc, double ang, double w)
{
    double l = acc + ang * w / 2;
    double r = acc - ang * w / 2;
    return (fabs(l) < fabs(r)) ? l : r;
}
__device__ __forceinline__ double max_speed_at_curvature(double V, double w, double ak)
{
    if (ak < 1e-6) return V;
    double m = ((2 * V / w) * V) / (ak * V + (2 * V / w));
    return pymin(m, V);
}

// generate_trapezoidal_profile (one_dim_mp_generator.py:4-69): closed form per sample
struct Trapezoid {
    double ttm, vmax, total_time, acc;
    long long K;
};
__device__ __forceinline__ Trapezoid trapezoid_setup(double max_velocity, double max_acceleration,
                                                     double total_distance, double time_step)
{
    Trapezoid T;
    double ttm = max_velocity / max_acceleration;
    double dist_accel = 0.5 * max_acceleration * (ttm * ttm);
    double total_time;
    if (2 * dist_accel > total_distance) {
        ttm = sqrt(total_distance / max_acceleration);
        max_velocity = max_acceleration * ttm;
        total_time = 2 * ttm;
    } else {
        double dc = total_distance - 2 * dist_accel;
        double tc = dc / max_velocity;
        total_time = 2 * ttm + tc;
    }
    double stop = total_time + time_step;
    double k = ceil((stop - 0.0) / time_step);    // len(np.arange(0, stop, step))
    T.K = (k > 0.0) ? ((k < 9.0e18) ? (long long)k : 9000000000000000000LL) : 0;
    T.ttm = ttm; T.vmax = max_velocity; T.total_time = total_time; T.acc = max_acceleration;
    return T;
}
__device__ __forceinline__ double trapezoid_vel(const Trapezoid& T, long long i, double time_step)
{
    double t = (double)i * time_step;
    if (t <= T.ttm) return T.acc * t;
    if (t <= T.total_time - T.ttm) return T.vmax;
    double tid = t - (T.total_time - T.ttm);
    return T.vmax - T.acc * tid;
}
