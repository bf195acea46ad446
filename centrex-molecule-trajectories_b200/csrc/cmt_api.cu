// cmt_api.cu -- the C ABI declared in include/cmt.h (libcmt_b200.so).
//
// Host side only: validation, flattening into the kernel-parameter table,
// exact threshold/slope precomputation, launches, staging for the host-buffer
// entry points.  No torch, no C++ types across the boundary.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <vector>

#include "cmt_kernels.cuh"

using namespace cmt;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t err__ = (expr);                                                         \
        if (err__ != cudaSuccess)                                                           \
            return fail(CMT_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                        __FILE__, __LINE__);                                                \
    } while (0)

// Entry points run on the handle's device but leave the caller's current device as they found it
// (PyTorch and other runtimes in the same process read it back with cudaGetDevice).
struct DeviceGuard {
    int prev = -1;
    cudaError_t status;
    explicit DeviceGuard(int device)
    {
        cudaGetDevice(&prev);
        status = prev == device ? cudaSuccess : cudaSetDevice(device);
        if (prev == device) prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

extern "C" const char *cmt_last_error(void) { return g_err; }

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

// kernel launches issued by this library (bench.py reports them)
static std::atomic<int64_t> g_launches{0};
static inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t cmt_launch_count(int reset)
{
    return reset ? g_launches.exchange(0) : g_launches.load();
}

static int g_debug_flags = 0;
extern "C" int cmt_debug_flags(int flags)
{
    const int old = g_debug_flags;
    if (flags >= 0) g_debug_flags = flags;
    return old;
}
extern "C" int cmt_version(void) { return CMT_VERSION; }

// ---------------------------------------------------------------------------
// beamline handle
// ---------------------------------------------------------------------------
struct PlaneD {
    double z;
    int kind, fate;
    double T;        // circle: squared-radius threshold
    double e[4];     // box: x1, x2, y1, y2
};

struct HostPipe;

struct cmt_beamline {
    Params P;
    int device;
    int max_rows;
    int n_sm;
    int math;           // CMT_MATH_EXACT / CMT_MATH_CONTRACTED
    double4 *d_tab;     // [tab_total]: (r_j, r_{j+1}, a_j, slope_j)
    size_t tab_bytes;   // dynamic shared memory the tail/trajectory kernels need (every table, plain layout)
    int seg_copies = 1; // lens_seg_kernel: replication of the first lens' table in shared memory (8, 4, 2 or 1)
    size_t seg_bytes = 0; // ... and the dynamic shared memory that takes
    bool has_mesh;      // a Honeycomb is present: launch the kernel variants that carry its hit test
    std::vector<struct PlaneD> planes;   // the filter planes in binary64 (thresholds of the quick filter derive from them)
    // staging of the host-buffer entry points: streams, device buffers, workspace.  Owned by the handle, so
    // handles on different devices (or two handles on one device) never share or thrash a pipe; calls on ONE
    // handle are serialised by pipe_mu.
    std::mutex pipe_mu;
    HostPipe *pipe = nullptr;
};

// Largest double s with sqrt(s) <= R under round-to-nearest, so that the
// reference's test `sqrt(x^2+y^2) > R` is exactly `x^2+y^2 > T`.
static double radius_threshold(double R)
{
    if (std::isnan(R)) return std::numeric_limits<double>::infinity();   // rho > NaN is never true
    if (R < 0) return -std::numeric_limits<double>::infinity();          // rho > R for every rho >= 0
    if (std::isinf(R)) return R;
    volatile double c = R * R;
    const double inf = std::numeric_limits<double>::infinity();
    for (int it = 0; it < 64 && std::sqrt((double)c) > R; ++it) c = std::nextafter((double)c, -inf);
    for (int it = 0; it < 64; ++it) {
        const double up = std::nextafter((double)c, inf);
        volatile double su = std::sqrt(up);
        if (su <= R) c = up; else break;
    }
    return c;
}

// ---------------------------------------------------------------------------
// Thresholds of the quick filter (cmt_device.cuh: quick_fate), in binary64, rounded outwards.
//
// Error model as in filter_fate, u = 2^-24: for a molecule with input errors e_* and 1/vz known to
// relv = e_vz/|vz| + 4u, the single-precision position at plane p is off by at most E/2 per coordinate,
//   E = c0 + c1 dtE + c2 dtE^2,  dtE = (|z_p| + |z_0|)/|vz|,
//   c0 = 2 (e_x0 + e_y0 + 2u (|x0| + |y0|)),  c1 = 2 ((|vx| + |vy|)(relv + 6u) + e_vx + e_vy),  c2 = 4 g (relv + 3u).
// quick_fate evaluates x = bx + sx z, y = A + B z + C z^2 with sx = vx/vz etc.; term by term its
// roundings are bounded by the same expression with 8u and 4u in place of 6u and 3u (slope: relv + u,
// intercept: one more rounding of |x0| + |sx z0|, plane position: u |z_p|; gravity: 2 relv + 8u on
// g (|z_p| + |z_0|)^2 / (2 vz^2)).  The guards pos_g .. ex_g bound every molecule-dependent factor:
//   c0 <= k0 = 2 (ex_g + 2u pos_g),  c1/|vz| <= k1 = 2 (ang_g (relv + 8u) + eva_g),  c2/vz^2 <= k2 = 4 g (relv + 4u) ainv_g^2,
// hence E <= eps_p = k0 + k1 Z_p + k2 Z_p^2 with Z_p = |z_p| + z0_g.
// Circle: |s32 - s| <= B(s32) = 2 (sqrt(2 s32 (1 + 4u)) + eps) eps + 8u s32 (twice the first-order bound, as in
// filter_fate, with |x| + |y| <= sqrt(2 (x^2 + y^2))).  B is increasing, so  s32 < T_lo := T - B(T)  implies s < T, and
// s32 > T_hi with T_hi - B(T_hi) >= T implies s > T (s - B(s) is increasing beyond 2 eps^2; T_hi >= 16 eps^2 is required).
// Box: the edges move by eps + 2^-21 max|edge| inwards (surely inside) and outwards (surely outside).
// Guard values: replayed inputs carry e_q = u |q|, so only pos_g, z0_g, ang_g, ainv_g actually select;
// for the Philox source the e_* guards are four times the typical error of draw_f32 (a molecule with a
// larger one is undecided, like one with (|vx| + |vy|) > 2 |vz| or slower than 8 sqrt(2 g Z)).
// ---------------------------------------------------------------------------
static float float_down(double x)
{
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
static float float_up(double x)
{
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}

static void build_quick(const std::vector<PlaneD> &planes, double g, const cmt_source_t *S, QuickFilter &Q)
{
    memset(&Q, 0, sizeof(Q));
    if (planes.empty()) return;
    const double u = std::ldexp(1.0, -24);
    const double inf = std::numeric_limits<double>::infinity();
    double lxy = 0, zmin = inf, zmax = 0;
    for (const PlaneD &p : planes) {
        zmax = std::max(zmax, std::fabs(p.z));
        zmin = std::min(zmin, std::fabs(p.z));
        if (p.kind == CMT_FILTER_CIRCLE) lxy = std::max(lxy, std::sqrt(p.T));
        else for (double v : p.e) if (std::isfinite(v)) lxy = std::max(lxy, std::fabs(v));
    }
    const double pos_g = 4 * lxy, ang_g = 2.0;
    double z0_g, ex_g, eva_g, relvz_g;
    if (!S) {
        z0_g = 2 * zmin;
        ex_g = 1.01 * u * pos_g; eva_g = 1.01 * u * ang_g; relvz_g = 1.01 * u;
    } else {
        const double en = std::ldexp(1.0, -18) * (2 * 1.25 + 0.8), k = 4.0, r22 = std::ldexp(1.0, -22);
        const double sxy = std::fabs(S->vsigma[0]) + std::fabs(S->vsigma[1]), sz = std::fabs(S->vsigma[2]);
        const double mz = std::fabs(S->vmean[2]);
        const double vref = std::max(mz - 2 * sz, 0.5 * mz);
        if (!(vref > 0) || !std::isfinite(vref) || !std::isfinite(sxy) || !std::isfinite(S->z)) return;
        z0_g = std::fabs(S->z) * (1 + 4 * u) + 1e-30;
        eva_g = k * (sxy * en + r22 * (std::fabs(S->vmean[0]) + std::fabs(S->vmean[1]) + 1.25 * sxy)) / vref;
        relvz_g = k * (sz * en + r22 * (mz + 1.25 * sz)) / vref;
        ex_g = S->pos_kind == CMT_POS_DISC ? k * 2 * std::ldexp(1.0, -18) * std::fabs(S->p0)
                                           : k * (std::fabs(S->p0) + std::fabs(S->p1)) * (en + 1.25 * r22);
    }
    const double zsum = zmax + z0_g;
    const double vmin = 8 * std::sqrt(2 * std::fabs(g) * zsum);
    const double ainv_g = vmin > 1e-12 ? 1 / vmin : 1e12;
    const double relv = relvz_g * 1.001 + 4 * u;
    const double k0 = 2 * (ex_g + 2 * u * pos_g), k1 = 2 * (ang_g * (relv + 8 * u) + eva_g);
    const double k2 = 4 * std::fabs(g) * (relv + 4 * u) * ainv_g * ainv_g;
    if (!(std::isfinite(k0) && std::isfinite(k1) && std::isfinite(k2))) return;
    for (size_t i = 0; i < planes.size(); ++i) {
        const PlaneD &p = planes[i];
        QuickPlane &q = Q.pl[i];
        const double Z = std::fabs(p.z) + z0_g;
        const double eps = (k0 + k1 * Z + k2 * Z * Z) * (1 + std::ldexp(1.0, -10));
        q.z = (float)p.z; q.kind = p.kind; q.fate = p.fate;
        if (p.kind == CMT_FILTER_CIRCLE) {
            auto B = [&](double s) { return 2 * (std::sqrt(2 * s * (1 + 4 * u)) + eps) * eps + 8 * u * s + 1e-37; };
            const double T_lo = p.T - B(p.T);
            double T_hi = p.T + B(p.T);
            for (int it = 0; it < 6; ++it) T_hi = p.T + B(T_hi);
            T_hi *= 1 + std::ldexp(1.0, -20);
            if (!(T_lo >= 0.5 * p.T) || !(T_hi >= 16 * eps * eps) || !(T_hi <= 2 * p.T)) return;
            q.v[0] = float_down(T_lo);
            q.v[1] = float_up(T_hi);
        } else {
            double m = 0;
            for (double v : p.e) if (std::isfinite(v)) m = std::max(m, std::fabs(v));
            const double eb = eps + std::ldexp(m, -21) + 1e-37;
            q.v[0] = float_up(p.e[0] + eb);   q.v[1] = float_down(p.e[1] - eb);
            q.v[2] = float_up(p.e[2] + eb);   q.v[3] = float_down(p.e[3] - eb);
            q.v[4] = float_down(p.e[0] - eb); q.v[5] = float_up(p.e[1] + eb);
            q.v[6] = float_down(p.e[2] - eb); q.v[7] = float_up(p.e[3] + eb);
            if (!(q.v[0] < q.v[1]) || !(q.v[2] < q.v[3])) return;
        }
    }
    Q.pos_g = float_down(pos_g); Q.z0_g = float_down(z0_g); Q.ainv_g = float_down(ainv_g); Q.ang_g = float_down(ang_g);
    Q.eva_g = float_down(eva_g); Q.relvz_g = float_down(relvz_g); Q.ex_g = float_down(ex_g);
    Q.usable = 1;
    Q.all_circles = 1;
    for (const PlaneD &p : planes) if (p.kind != CMT_FILTER_CIRCLE) Q.all_circles = 0;
}

// May the straight-line force evaluation (cmt_device.cuh: lens_acc_fast) be used with this table?  Its validity
// record relies on: every knot, value and slope at most 2^400 in magnitude (no quotient overflows, no divisor
// reaches 2^1017); knots non-negative; and every interval width r_{j+1} - r_j and every difference r - r_j for r in
// [r_j, r_{j+1}) exact in binary64, which holds when r_{j+1} <= 2 r_j (Sterbenz) or r_j = 0.  The evenly spaced
// tables the reference builds (linspace from 0) qualify; any other table is evaluated on the reference path.
static bool table_fast_ok(const cmt_table_t &tb)
{
    const double big = std::ldexp(1.0, 400);
    for (int i = 0; i < tb.n; ++i) {
        if (!(tb.r[i] >= 0.0) || !(tb.r[i] <= big) || !(std::fabs(tb.a[i]) <= big)) return false;
        if (i + 1 < tb.n) {
            if (!(tb.r[i] == 0.0 || tb.r[i + 1] <= 2.0 * tb.r[i])) return false;
            volatile double num = tb.a[i + 1] - tb.a[i];
            volatile double den = tb.r[i + 1] - tb.r[i];
            const double slope = num / den;
            if (!(std::fabs(slope) <= big)) return false;
        }
    }
    return true;
}

extern "C" int cmt_beamline_create(const cmt_element_t *elements, int n_elements, const cmt_table_t *tables,
                                   int n_tables, int n_fates, int fate_detected, double g, int device,
                                   cmt_beamline_t **out)
{
    if (!out) return fail(CMT_EINVAL, "out is NULL");
    *out = nullptr;
    if (n_elements < 0 || n_elements > CMT_MAX_ELEMENTS)
        return fail(CMT_EINVAL, "n_elements=%d outside [0,%d]", n_elements, CMT_MAX_ELEMENTS);
    if (n_elements > 0 && !elements) return fail(CMT_EINVAL, "elements is NULL");
    if (n_fates < 1 || n_fates > CMT_MAX_FATES)
        return fail(CMT_EINVAL, "n_fates=%d outside [1,%d]", n_fates, CMT_MAX_FATES);
    if (fate_detected < 0 || fate_detected >= n_fates) return fail(CMT_EINVAL, "fate_detected out of range");
    if (n_tables < 0 || n_tables > CMT_MAX_TABLES)
        return fail(CMT_EINVAL, "n_tables=%d outside [0,%d]", n_tables, CMT_MAX_TABLES);

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(CMT_ENODEV, "no CUDA device visible");
    }
    if (device < 0 || device >= n_dev) return fail(CMT_EINVAL, "device %d not in [0,%d)", device, n_dev);
    // compute capability and SM count, asked once per device: cudaGetDeviceProperties takes milliseconds, and a
    // sweep creates a handle per point (40 handles: 0.11 s of a 0.22 s run_sweep call)
    struct DevInfo { int major = 0, minor = 0, n_sm = 0; };
    static std::mutex dev_mu;
    static DevInfo dev_info[64];
    DevInfo prop;
    {
        std::lock_guard<std::mutex> lk(dev_mu);
        DevInfo &slot = dev_info[device < 64 ? device : 63];
        if (device >= 63 || slot.n_sm == 0) {
            DevInfo q;
            CUDA_TRY(cudaDeviceGetAttribute(&q.major, cudaDevAttrComputeCapabilityMajor, device));
            CUDA_TRY(cudaDeviceGetAttribute(&q.minor, cudaDevAttrComputeCapabilityMinor, device));
            CUDA_TRY(cudaDeviceGetAttribute(&q.n_sm, cudaDevAttrMultiProcessorCount, device));
            if (device < 63) slot = q;
            prop = q;
        } else {
            prop = slot;
        }
    }
    if (prop.major != 10)
        return fail(CMT_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);

    std::vector<int> tab_off(std::max(n_tables, 1), 0);
    int tab_total = 0;
    for (int t = 0; t < n_tables; ++t) {
        if (!tables[t].r || !tables[t].a || tables[t].n < 2)
            return fail(CMT_EINVAL, "lens table %d needs >= 2 points", t);
        for (int i = 1; i < tables[t].n; ++i)
            if (!(tables[t].r[i] > tables[t].r[i - 1]))
                return fail(CMT_EINVAL, "lens table %d: r must be strictly ascending (index %d)", t, i);
        tab_off[t] = tab_total;
        tab_total += tables[t].n;
    }

    cmt_beamline *bl = new cmt_beamline();
    memset(&bl->P, 0, sizeof(bl->P));
    Params &P = bl->P;
    P.n_el = n_elements;
    P.n_fates = n_fates;
    P.fate_detected = fate_detected;
    P.first_lens = n_elements;
    P.g = g;
    P.tab_total = tab_total;
    P.flags = g_debug_flags;
    bl->device = device;
    bl->n_sm = prop.n_sm;
    bl->max_rows = 1;
    bl->math = CMT_MATH_EXACT;
    bl->d_tab = nullptr;
    bl->has_mesh = false;

    for (int i = 0; i < n_elements; ++i) {
        const cmt_element_t &s = elements[i];
        DevElement &d = P.el[i];
        if (i > 0 && s.z0 < elements[i - 1].z0) {
            delete bl;
            return fail(CMT_EINVAL, "elements must be sorted by z0 (element %d)", i);
        }
        if (s.fate < 0 || s.fate >= n_fates) { delete bl; return fail(CMT_EINVAL, "element %d: fate out of range", i); }
        d.type = s.type; d.fate = s.fate; d.fate2 = s.fate2; d.n_steps = s.n_steps;
        d.z0 = s.z0; d.z1 = s.z1;
        switch (s.type) {
        case CMT_CIRCULAR:
            d.p[0] = radius_threshold(s.R);
            bl->max_rows += 2;
            break;
        case CMT_RECTANGULAR:
            d.p[0] = s.x1; d.p[1] = s.x2; d.p[2] = s.y1; d.p[3] = s.y2;
            bl->max_rows += 2;
            break;
        case CMT_FIELDPLATES:
            d.p[0] = s.x1; d.p[1] = s.x2;
            bl->max_rows += 2;
            break;
        case CMT_LENS: {
            if (s.table < 0 || s.table >= n_tables) { delete bl; return fail(CMT_EINVAL, "element %d: lens table index out of range", i); }
            if (s.fate2 < 0 || s.fate2 >= n_fates) { delete bl; return fail(CMT_EINVAL, "element %d: fate2 out of range", i); }
            if (s.n_steps < 0) { delete bl; return fail(CMT_EINVAL, "element %d: n_steps < 0", i); }
            const cmt_table_t &tb = tables[s.table];
            d.p[0] = radius_threshold(s.R);
            d.p[1] = s.dz;
            d.p[2] = (tb.n - 1) / (tb.r[tb.n - 1] - tb.r[0]);
            d.p[3] = table_fast_ok(tb) ? 1.0 : 0.0;
            d.tab_off = tab_off[s.table];
            d.tab_len = tb.n;
            bl->max_rows += 2 + s.n_steps;
            if (P.first_lens == n_elements) P.first_lens = i;
            break;
        }
        case CMT_HONEYCOMB:
            if (s.n_steps < 1 || s.reserved < 1 || (int64_t)s.n_steps * s.reserved > (1 << 30)) {
                delete bl;
                return fail(CMT_EINVAL, "element %d: honeycomb needs nx (n_steps) >= 1 and ny (reserved) >= 1", i);
            }
            if (!(s.dz > 0.0)) { delete bl; return fail(CMT_EINVAL, "element %d: honeycomb pitch (dz) must be > 0", i); }
            d.p[0] = s.R; d.p[1] = s.dz; d.p[2] = s.x1; d.p[3] = s.y1;
            d.tab_len = s.reserved;
            bl->has_mesh = true;
            bl->max_rows += 2;
            break;
        default:
            delete bl;
            return fail(CMT_EINVAL, "element %d: unknown type %d", i, s.type);
        }
    }

    // leading run of circular planes for the walk kernel's tight loop
    {
        FastPlanes &F = P.fast;
        F.n = 0; F.next_element = 0; F.ends_at_lens = 0;
        int e = 0;
        while (e < n_elements && P.el[e].type == CMT_CIRCULAR && F.n + 2 <= CMT_MAX_FAST_PLANES) {
            for (int k = 0; k < 2; ++k) {
                F.z[F.n] = k == 0 ? P.el[e].z0 : P.el[e].z1;
                F.T[F.n] = P.el[e].p[0];
                F.fate[F.n] = P.el[e].fate;
                ++F.n;
            }
            ++e;
        }
        if (e < n_elements && e == P.first_lens && F.n + 1 <= CMT_MAX_FAST_PLANES) {
            F.z[F.n] = P.el[e].z0;
            F.T[F.n] = P.el[e].p[0];
            F.fate[F.n] = P.el[e].fate;     // "Lens entrance"
            ++F.n;
            F.ends_at_lens = 1;
            ++e;
        }
        F.next_element = e;
    }

    // FP32 fate filter: the leading run of circular / rectangular / field-plate elements as
    // single-precision planes, closed by the first lens' entrance plane when it follows directly.
    // Built only when every threshold is an ordinary number (the error model of filter_fate
    // assumes normal single-precision magnitudes); otherwise n = 0 and the walk is binary64 only.
    {
        FilterPlanes &F = P.filt;
        F.n = 0; F.covers_all = 0;
        F.hg = (float)(0.5 * g);
        F.g_abs = (float)std::fabs(g);
        bool usable = std::isfinite(g) && std::fabs(g) < 1e6;
        const float inf = std::numeric_limits<float>::infinity();
        auto ordinary = [](double v) { return std::isfinite(v) && std::fabs(v) < 1e15; };
        auto edge_ok = [](double v) { return !std::isnan(v) && (std::isinf(v) || std::fabs(v) < 1e15); };
        auto circle = [&](double z, double T, int fate) {
            if (!(ordinary(z) && std::isfinite(T) && T >= 1e-30 && T <= 1e30)) { usable = false; return; }
            FilterPlane &pl = F.pl[F.n++];
            pl.z = (float)z; pl.kind = CMT_FILTER_CIRCLE; pl.fate = fate;
            pl.a = (float)T; pl.b = pl.c = pl.d = 0.f;
            pl.tol = std::ldexp(std::fabs(pl.a), -21) + 1e-37f;
            bl->planes.push_back(PlaneD{z, CMT_FILTER_CIRCLE, fate, T, {0, 0, 0, 0}});
        };
        auto box = [&](double z, double x1, double x2, double y1, double y2, int fate) {
            if (!(ordinary(z) && edge_ok(x1) && edge_ok(x2) && edge_ok(y1) && edge_ok(y2))) { usable = false; return; }
            FilterPlane &pl = F.pl[F.n++];
            pl.z = (float)z; pl.kind = CMT_FILTER_BOX; pl.fate = fate;
            pl.a = (float)x1; pl.b = (float)x2; pl.c = (float)y1; pl.d = (float)y2;
            float m = 0.f;
            for (float v : {pl.a, pl.b, pl.c, pl.d}) if (std::isfinite(v)) m = std::max(m, std::fabs(v));
            pl.tol = std::ldexp(m, -21) + 1e-37f;
            bl->planes.push_back(PlaneD{z, CMT_FILTER_BOX, fate, 0.0, {x1, x2, y1, y2}});
        };
        int e = 0;
        for (; usable && e < n_elements && F.n + 2 <= CMT_MAX_FILTER_PLANES; ++e) {
            const DevElement &d = P.el[e];
            if (d.type == CMT_CIRCULAR) {
                circle(d.z0, d.p[0], d.fate);
                if (usable) circle(d.z1, d.p[0], d.fate);
            } else if (d.type == CMT_RECTANGULAR) {
                box(d.z0, d.p[0], d.p[1], d.p[2], d.p[3], d.fate);
                if (usable) box(d.z1, d.p[0], d.p[1], d.p[2], d.p[3], d.fate);
            } else if (d.type == CMT_FIELDPLATES) {
                // x only, at z0 and (look-ahead or wall crossing, same fate and one row either way) at z1
                box(d.z0, d.p[0], d.p[1], -inf, inf, d.fate);
                if (usable) box(d.z1, d.p[0], d.p[1], -inf, inf, d.fate);
            } else {
                break;
            }
        }
        if (usable && e < n_elements && e == P.first_lens && F.n + 1 <= CMT_MAX_FILTER_PLANES)
            circle(P.el[e].z0, P.el[e].p[0], P.el[e].fate);      // "Lens entrance"
        else if (usable && e == n_elements)
            F.covers_all = 1;
        if (!usable) { F.n = 0; F.covers_all = 0; bl->planes.clear(); }
        build_quick(bl->planes, g, nullptr, P.quick);          // replayed initial conditions; Philox launches rebuild it
    }

    bl->tab_bytes = (size_t)tab_total * sizeof(double4);
    if (tab_total > 0) {
        std::vector<double4> h((size_t)tab_total);
        for (int t = 0; t < n_tables; ++t) {
            const cmt_table_t &tb = tables[t];
            for (int i = 0; i < tb.n; ++i) {
                double4 &e = h[tab_off[t] + i];
                e.x = tb.r[i];
                e.y = -std::numeric_limits<double>::infinity();
                e.z = tb.a[i];
                e.w = 0.0;
                if (i + 1 < tb.n) {
                    // np.interp: slope = (fp[j+1]-fp[j]) / (xp[j+1]-xp[j]); same IEEE ops here
                    volatile double num = tb.a[i + 1] - tb.a[i];
                    volatile double den = tb.r[i + 1] - tb.r[i];
                    e.y = den;                 // interval width (exact for tables that pass table_fast_ok)
                    e.w = num / den;
                }
            }
        }
        if (bl->tab_bytes > 200 * 1024) { delete bl; return fail(CMT_EINVAL, "lens tables too large for shared memory (%zu B)", bl->tab_bytes); }
        DeviceGuard guard(device);
        cudaError_t e = guard.status;
        if (e == cudaSuccess) e = cudaMalloc(&bl->d_tab, bl->tab_bytes);
        if (e == cudaSuccess) e = cudaMemcpy(bl->d_tab, h.data(), bl->tab_bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            if (bl->d_tab) cudaFree(bl->d_tab);
            delete bl;
            return fail(CMT_ECUDA, "uploading lens tables failed: %s", cudaGetErrorString(e));
        }
        P.tab = bl->d_tab;
        // lens_seg_kernel keeps only the first lens' table, replicated so that the lanes of a quarter warp read
        // different bank groups (Table, cmt_device.cuh): the most copies that still leave every CTA the registers
        // admit (LENS_SEG_MIN_CTAS per SM) its shared memory (227 KB per SM, 1 KB reserved per CTA).  Measured
        // (profiles/README.md, round 2, 222-point table): 1 / 2 / 4 / 8 copies -> lens stage 0.495 / 0.473 / 0.468 /
        // 0.472 ms alone at 1e7 molecules, overlapped step 0.365 / 0.350 / 0.345 / 0.352 ms.
        bl->seg_copies = 1;
        bl->seg_bytes = 0;
        if (P.first_lens < P.n_el) {
            static const int tune_copies = env_int("CMT_TUNE_SEG_COPIES", 0);            // experiments only
            const size_t one = (size_t)P.el[P.first_lens].tab_len * sizeof(double4);
            const size_t budget = (size_t)(227 * 1024) / LENS_SEG_MIN_CTAS - 2048;
            int copies = 8;
            while (copies > 1 && one * copies > budget) copies >>= 1;
            if (tune_copies == 1 || tune_copies == 2 || tune_copies == 4 || tune_copies == 8)
                if (one * tune_copies <= 200 * 1024) copies = tune_copies;
            bl->seg_copies = copies;
            bl->seg_bytes = one * copies;
        }
        if (bl->tab_bytes > 40 * 1024 || bl->seg_bytes > 40 * 1024) {
            // Tables beyond the default 48 KB of dynamic shared memory need the opt-in.  The attribute is per
            // kernel and per device and must cover EVERY live handle, so it is only ever raised.
            static std::mutex mu;
            static size_t granted[64] = {0};
            std::lock_guard<std::mutex> lk(mu);
            const size_t need = std::max(bl->tab_bytes, bl->seg_bytes);
            if (device < 64 && need > granted[device]) {
                const int b = (int)need;
                cudaError_t a = cudaSuccess;
                auto opt_in = [&](const void *fn) { if (a == cudaSuccess) a = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, b); };
                opt_in((const void *)lens_seg_kernel<false, 1>); opt_in((const void *)lens_seg_kernel<true, 1>);
                opt_in((const void *)lens_seg_kernel<false, 2>); opt_in((const void *)lens_seg_kernel<true, 2>);
                opt_in((const void *)lens_seg_kernel<false, 4>); opt_in((const void *)lens_seg_kernel<true, 4>);
                opt_in((const void *)lens_seg_kernel<false, 8>); opt_in((const void *)lens_seg_kernel<true, 8>);
                opt_in((const void *)tail_kernel<false, false>); opt_in((const void *)tail_kernel<false, true>);
                opt_in((const void *)tail_kernel<true, false>);  opt_in((const void *)tail_kernel<true, true>);
                opt_in((const void *)trajectory_kernel<false>); opt_in((const void *)trajectory_kernel<true>);
                opt_in((const void *)crossing_kernel<false>);   opt_in((const void *)crossing_kernel<true>);
                if (a != cudaSuccess) {
                    cudaFree(bl->d_tab);
                    delete bl;
                    return fail(CMT_ECUDA, "shared-memory opt-in for %d B of lens tables failed: %s", b, cudaGetErrorString(a));
                }
                granted[device] = need;
            }
        }
    }
    *out = bl;
    return CMT_OK;
}

static void pipe_destroy(HostPipe *p);

extern "C" void cmt_beamline_destroy(cmt_beamline_t *bl)
{
    if (!bl) return;
    pipe_destroy(bl->pipe);
    if (bl->d_tab) {
        DeviceGuard guard(bl->device);
        cudaFree(bl->d_tab);
    }
    delete bl;
}

extern "C" int cmt_beamline_set_math(cmt_beamline_t *bl, int mode)
{
    if (!bl) return fail(CMT_EINVAL, "beamline handle is NULL");
    if (mode != CMT_MATH_EXACT && mode != CMT_MATH_CONTRACTED) return fail(CMT_EINVAL, "unknown math mode %d", mode);
    bl->math = mode;
    return CMT_OK;
}

extern "C" int cmt_beamline_max_rows(const cmt_beamline_t *bl) { return bl ? bl->max_rows : CMT_EINVAL; }
extern "C" int cmt_beamline_device(const cmt_beamline_t *bl) { return bl ? bl->device : CMT_EINVAL; }

static constexpr size_t WS_HEADER = 256;

extern "C" size_t cmt_workspace_bytes(const cmt_beamline_t *bl, int64_t n_max)
{
    if (!bl || n_max < 0) return 0;
    if (bl->P.first_lens >= bl->P.n_el) return WS_HEADER;
    // two queue arrays that alternate between the lens segments: the walk kernel fills the first one with
    // the survivors of the lens entrance test; molecules leave the lens only in the last segment, which has
    // no successor, so the exit queue takes the array that segment would have written survivors to
    return WS_HEADER + 2 * (size_t)QUEUE_COMPONENTS * sizeof(double) * (size_t)n_max;
}

// ---------------------------------------------------------------------------
// timing (CUDA events on the launch stream)
// ---------------------------------------------------------------------------
struct TimedLaunch {
    cudaEvent_t a, b;
    int kind;
    cudaStream_t st;
};
static std::mutex g_time_mu;
static int g_time_level = 0;      // 0 off; 1 the stage timers (kinds 0..3); 2 also one timer per lens kernel (kinds 8..)
static std::vector<TimedLaunch> g_time_pending;
static double g_time_ms[4] = {0, 0, 0, 0};
static int64_t g_time_n[4] = {0, 0, 0, 0};

// kinds: 0 walk kernel, 1 lens stage (segment launches + tail), 2 trajectory / resume kernels, 3 reserved;
// timeline only (level 2): 8 + k = lens segment k, 7 = tail kernel
struct ScopedTimer {
    bool on;
    cudaEvent_t a, b;
    cudaStream_t st;
    int kind;
    ScopedTimer(int kind_, cudaStream_t st_, int level = 1) : on(false), st(st_), kind(kind_)
    {
        std::lock_guard<std::mutex> lk(g_time_mu);
        on = g_time_level >= level;
        if (on) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
    }
    ~ScopedTimer()
    {
        if (on) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lk(g_time_mu);
            g_time_pending.push_back({a, b, kind, st});
        }
    }
};

extern "C" int cmt_timing_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_time_mu);
    g_time_level = on < 0 ? 0 : on;
    return CMT_OK;
}

static void timing_fold(const TimedLaunch &t)
{
    float f = 0;
    if (t.kind < 4 && cudaEventElapsedTime(&f, t.a, t.b) == cudaSuccess) {
        g_time_ms[t.kind] += f;
        g_time_n[t.kind] += 1;
    }
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
}

extern "C" int cmt_timing_read(double ms[4], int64_t launches[4], int reset)
{
    std::lock_guard<std::mutex> lk(g_time_mu);
    for (auto &t : g_time_pending) {
        cudaEventSynchronize(t.b);
        timing_fold(t);
    }
    g_time_pending.clear();
    for (int k = 0; k < 4; ++k) {
        if (ms) ms[k] = g_time_ms[k];
        if (launches) launches[k] = g_time_n[k];
        if (reset) { g_time_ms[k] = 0; g_time_n[k] = 0; }
    }
    return CMT_OK;
}

extern "C" int64_t cmt_timing_timeline(double *start_ms, double *end_ms, int32_t *kind, int32_t *stream_id, int64_t capacity)
{
    std::lock_guard<std::mutex> lk(g_time_mu);
    if (g_time_pending.empty()) return 0;
    for (auto &t : g_time_pending) cudaEventSynchronize(t.b);
    // the earliest start is the origin: elapsed times against the first record, shifted by the minimum
    const cudaEvent_t ref = g_time_pending.front().a;
    std::vector<double> a(g_time_pending.size()), b(g_time_pending.size());
    std::vector<cudaStream_t> streams;
    double lo = 0.0;
    for (size_t i = 0; i < g_time_pending.size(); ++i) {
        float fa = 0, fb = 0;
        cudaEventElapsedTime(&fa, ref, g_time_pending[i].a);
        cudaEventElapsedTime(&fb, ref, g_time_pending[i].b);
        a[i] = fa; b[i] = fb;
        lo = std::min(lo, (double)fa);
    }
    int64_t n = 0;
    for (size_t i = 0; i < g_time_pending.size(); ++i) {
        const TimedLaunch &t = g_time_pending[i];
        size_t sid = 0;
        while (sid < streams.size() && streams[sid] != t.st) ++sid;
        if (sid == streams.size()) streams.push_back(t.st);
        if (n < capacity) {
            if (start_ms) start_ms[n] = a[i] - lo;
            if (end_ms) end_ms[n] = b[i] - lo;
            if (kind) kind[n] = t.kind;
            if (stream_id) stream_id[n] = (int32_t)sid;
        }
        ++n;
        timing_fold(t);
    }
    g_time_pending.clear();
    return n;
}

// ---------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------
static int check_outputs(const cmt_beamline_t *bl, const cmt_outputs_t *out, int64_t n)
{
    if (!bl) return fail(CMT_EINVAL, "beamline handle is NULL");
    if (!out) return fail(CMT_EINVAL, "outputs is NULL");
    if (!out->counters) return fail(CMT_EINVAL, "outputs.counters is required");
    if (out->final_state && out->final_ld < n) return fail(CMT_EINVAL, "outputs.final_ld < n");
    if (out->saved_index && (!out->saved_count || out->saved_capacity < 0))
        return fail(CMT_EINVAL, "outputs.saved_index needs saved_count and a capacity");
    if (out->queue_capacity < 0) return fail(CMT_EINVAL, "outputs.queue_capacity < 0");
    if (out->queue_capacity > 0 && !out->work)
        return fail(CMT_EINVAL, "outputs.queue_capacity needs outputs.work (work[6] reports dropped molecules)");
    return CMT_OK;
}

static int propagate(const cmt_beamline_t *bl, bool philox, const cmt_source_t *src, uint64_t seed,
                     const double *ic, int64_t ic_ld, int64_t n, int64_t first_index,
                     const cmt_outputs_t *out, void *workspace, size_t workspace_bytes, void *stream)
{
    int rc = check_outputs(bl, out, n);
    if (rc) return rc;
    if (n < 0) return fail(CMT_EINVAL, "n < 0");
    if (n == 0) return CMT_OK;
    if (n >= (1ll << SEG_INDEX_BITS)) return fail(CMT_EINVAL, "n must be below 2^%d per launch", SEG_INDEX_BITS);
    const bool has_lens = bl->P.first_lens < bl->P.n_el;
    const int64_t q_cap = out->queue_capacity > 0 ? std::min<int64_t>(out->queue_capacity, n) : n;
    const size_t need = cmt_workspace_bytes(bl, q_cap);
    if (!workspace || workspace_bytes < need)
        return fail(CMT_ENOMEM, "workspace of %zu B given, %zu B needed for n=%lld with a lens queue of %lld", workspace_bytes,
                    need, (long long)n, (long long)q_cap);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return fail(CMT_EINVAL, "workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(bl->device);
    CUDA_TRY(guard.status);

    Queue Q;
    Q.count = reinterpret_cast<unsigned long long *>(workspace);
    Q.cursor = Q.count + 1;
    Q.q = reinterpret_cast<double *>(static_cast<char *>(workspace) + WS_HEADER);
    Q.cap = has_lens ? q_cap : 0;
    // header words: [0],[1] entry queue (count, cursor); [2],[3] exit queue; [4+2k],[5+2k] output of lens segment k
    Queue X;
    X.count = Q.count + 2;
    X.cursor = Q.count + 3;
    X.q = nullptr;               // set with the segment count below
    X.cap = Q.cap;
    CUDA_TRY(cudaMemsetAsync(workspace, 0, WS_HEADER, st));

    cmt_source_t S;
    memset(&S, 0, sizeof(S));
    if (src) S = *src;

    // pair mode of the walk kernel (two molecules per thread and turn): all-circular front end with constant
    // thresholds, Counter within the per-lane columns, no saved-index list, and -- for replayed initial
    // conditions -- components that can be read as aligned 16-byte pairs
    const bool pair_mode = bl->P.filt.n > 0 && bl->P.quick.usable && bl->P.quick.all_circles && !out->final_state &&
                           !out->saved_index && bl->P.n_fates <= PAIR_MAX_FATES &&
                           !(bl->P.flags & (CMT_FLAG_NO_FILTER | CMT_FLAG_NO_QUICK | CMT_FLAG_NO_PAIRS)) &&
                           (philox || ((reinterpret_cast<uintptr_t>(ic) & 15u) == 0 && (ic_ld & 1) == 0));
    const int per_tile = pair_mode ? 2 * WALK_THREADS : WALK_THREADS;
    const int64_t tiles = (n + per_tile - 1) / per_tile;
    static const int tune_walk_ctas = env_int("CMT_TUNE_WALK_CTAS", 0);   // experiments only
    static const int tune_lens_prio = env_int("CMT_TUNE_LENS_PRIO", 1);
    {
        ScopedTimer tm(0, st);
        // variants: source (replay / Philox) x arithmetic (exact / contracted) x Honeycomb test compiled in
        using WalkFn = void (*)(const Params, const cmt_source_t, uint64_t, const double *, int64_t, int64_t, int64_t,
                                const cmt_outputs_t, Queue, int);
        static const WalkFn walk[2][2][2] = {
            {{walk_kernel<false, false, false>, walk_kernel<false, false, true>},
             {walk_kernel<false, true, false>, walk_kernel<false, true, true>}},
            {{walk_kernel<true, false, false>, walk_kernel<true, false, true>},
             {walk_kernel<true, true, false>, walk_kernel<true, true, true>}}};
        const bool contract = bl->math == CMT_MATH_CONTRACTED;
        // Persistent grid: as many CTAs per SM as the variant's registers allow (one wave, every CTA loops over
        // tiles), capped at 14 for the one-molecule-per-thread replay form, whose prefetch was tuned there.
        const WalkFn fn = walk[philox][contract][bl->has_mesh];
        int walk_ctas = tune_walk_ctas;
        if (walk_ctas <= 0) {
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&walk_ctas, fn, WALK_THREADS, 0));
            if (!philox && !pair_mode) walk_ctas = std::min(walk_ctas, 14);
            // Philox launches (18 CTAs per SM would fit): beside a lens stage -- the other half of the same run --
            // eight walk CTAs per SM do as well as a full wave and leave the integrator's CTAs their registers: lone
            // run of 1e7 molecules 0.618 -> 0.584 ms (profiles/README.md, round 2).  Replay launches keep their full
            // wave of 10: alone the kernel is HBM-bound (0.091 against 0.098 ms), overlapped it makes no difference.
            if (has_lens && philox) walk_ctas = std::min(walk_ctas, 8);
            walk_ctas = std::max(walk_ctas, 1);
        }
        const int grid_walk = (int)std::min<int64_t>(tiles, (int64_t)bl->n_sm * walk_ctas);
        cudaLaunchConfig_t wcfg;
        memset(&wcfg, 0, sizeof(wcfg));
        wcfg.gridDim = dim3(grid_walk); wcfg.blockDim = dim3(WALK_THREADS);
        wcfg.stream = st;
        if (philox && bl->P.filt.n > 0) {
            // the quick filter's thresholds depend on the source's error bounds: per launch, by value
            Params P = bl->P;
            build_quick(bl->planes, bl->P.g, &S, P.quick);
            // the constant thresholds of a Philox launch must again be all-circular and usable for pair mode
            const int pm = pair_mode && P.quick.usable && P.quick.all_circles;
            CUDA_TRY(cudaLaunchKernelEx(&wcfg, walk[1][contract][bl->has_mesh], P, S, seed, (const double *)nullptr,
                                        (int64_t)0, n, first_index, *out, Q, pm));
        } else {
            CUDA_TRY(cudaLaunchKernelEx(&wcfg, walk[philox][contract][bl->has_mesh], bl->P, S, seed,
                                        philox ? (const double *)nullptr : ic, philox ? (int64_t)0 : ic_ld, n,
                                        first_index, *out, Q, (int)pair_mode));
        }
        count_launch();
    }
    CUDA_TRY(cudaGetLastError());
    if (has_lens) {
        // Lens stage: the first lens' integrator as a chain of segment launches, then the elements behind
        // it for the molecules that got through.  The integrator launches run at the highest stream
        // priority: when steps overlap on several streams, their CTAs are placed ahead of the next
        // step's walk CTAs, which fill what is left (measured: 0.549 -> 0.527 ms per step).
        ScopedTimer tm(1, st);
        static const int tune_seg = env_int("CMT_TUNE_SEG", LENS_SEGMENT_STEPS);          // experiments only
        // (a lone launch of 8e7 molecules would like 4 CTAs per SM -- FP64 pipe 75 % against 68 % -- but the 2^26-molecule
        // chunks of a large run overlap with their neighbours on the other streams like small ones do: 1e10 molecules
        // through run_simulation take 0.406 / 0.403 / 0.424 s with 2 / 3 / 4, profiles/experiments/r02x_*.log)
        static const int seg_per_sm = env_int("CMT_TUNE_SEG_CTAS", LENS_SEG_GRID_CTAS);   // experiments only
        const bool contract = bl->math == CMT_MATH_CONTRACTED;
        int least = 0, greatest = 0;
        if (tune_lens_prio) cudaDeviceGetStreamPriorityRange(&least, &greatest);
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributePriority;
        attr.val.priority = tune_lens_prio ? greatest : 0;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.blockDim = dim3(LENS_THREADS);
        cfg.dynamicSmemBytes = bl->seg_bytes; cfg.stream = st; cfg.attrs = &attr; cfg.numAttrs = 1;

        const int n_steps = bl->P.el[bl->P.first_lens].n_steps;
        constexpr int max_seg = (int)(WS_HEADER / 16) - 2;                          // a (count, cursor) pair per segment
        int seg = std::max(1, tune_seg);
        if (n_steps >= (1 << (63 - SEG_INDEX_BITS))) seg = n_steps;                 // the step count would not fit its bit field
        else if ((n_steps + seg - 1) / seg > max_seg) seg = (n_steps + max_seg - 1) / max_seg;
        int n_seg = std::max(1, (n_steps + seg - 1) / seg);
        // experiments only: CMT_TUNE_SEG_PLAN="75,75,150,300" gives every launch its own number of steps (the last entry
        // repeats until the lens is through)
        static const std::vector<int> tune_plan = [] {
            std::vector<int> v;
            if (const char *e = getenv("CMT_TUNE_SEG_PLAN")) {
                for (const char *p = e; *p;) {
                    char *end = nullptr;
                    const long k = strtol(p, &end, 10);
                    if (end == p) break;
                    if (k > 0) v.push_back((int)k);
                    p = *end ? end + 1 : end;
                }
            }
            return v;
        }();
        std::vector<int> plan;
        if (!tune_plan.empty() && n_steps < (1 << (63 - SEG_INDEX_BITS))) {
            for (int done = 0, k = 0; done < n_steps; ++k) {
                const int len = tune_plan[std::min<size_t>(k, tune_plan.size() - 1)];
                plan.push_back(len);
                done += len;
            }
            if ((int)plan.size() <= max_seg) n_seg = (int)plan.size(); else plan.clear();
        }
        const int64_t max_ctas = (n + LENS_THREADS - 1) / LENS_THREADS;
        cfg.gridDim = dim3((unsigned)std::min<int64_t>(max_ctas, (int64_t)bl->n_sm * seg_per_sm));
        X.q = Q.q + (size_t)(n_seg & 1) * QUEUE_COMPONENTS * (size_t)Q.cap;   // the last segment's unused output array
        using SegFn = void (*)(const Params, int64_t, const cmt_outputs_t, Queue, Queue, Queue, int);
        static const SegFn seg_fn[2][4] = {
            {lens_seg_kernel<false, 1>, lens_seg_kernel<false, 2>, lens_seg_kernel<false, 4>, lens_seg_kernel<false, 8>},
            {lens_seg_kernel<true, 1>, lens_seg_kernel<true, 2>, lens_seg_kernel<true, 4>, lens_seg_kernel<true, 8>}};
        const int copies_log2 = bl->seg_copies == 8 ? 3 : bl->seg_copies == 4 ? 2 : bl->seg_copies == 2 ? 1 : 0;
        for (int k = 0; k < n_seg; ++k) {
            Queue A, B;
            A.cap = B.cap = Q.cap;
            A.q = Q.q + (size_t)(k & 1) * QUEUE_COMPONENTS * (size_t)Q.cap;
            B.q = Q.q + (size_t)((k + 1) & 1) * QUEUE_COMPONENTS * (size_t)Q.cap;
            A.count = k == 0 ? Q.count : Q.count + 4 + 2 * (k - 1);
            A.cursor = A.count + 1;
            B.count = Q.count + 4 + 2 * k;
            B.cursor = B.count + 1;
            {
                ScopedTimer tk(8 + k, st, 2);
                CUDA_TRY(cudaLaunchKernelEx(&cfg, seg_fn[contract][copies_log2], bl->P, first_index, *out, A, B, X,
                                            plan.empty() ? seg : plan[k]));
            }
            count_launch();
        }
        const int grid_tail = (int)std::min<int64_t>((n + TRAJ_THREADS - 1) / TRAJ_THREADS, (int64_t)bl->n_sm * 8);
        using TailFn = void (*)(const Params, int64_t, const cmt_outputs_t, Queue);
        static const TailFn tail[2][2] = {{tail_kernel<false, false>, tail_kernel<false, true>},
                                          {tail_kernel<true, false>, tail_kernel<true, true>}};
        {
            ScopedTimer tk(7, st, 2);
            tail[contract][bl->has_mesh]<<<grid_tail, TRAJ_THREADS, bl->tab_bytes, st>>>(bl->P, first_index, *out, X);
        }
        count_launch();
    }
    CUDA_TRY(cudaGetLastError());
    return CMT_OK;
}

extern "C" int cmt_propagate_ic(const cmt_beamline_t *bl, int64_t n, int64_t first_index, const double *ic,
                                int64_t ic_ld, const cmt_outputs_t *out, void *workspace,
                                size_t workspace_bytes, void *stream)
{
    if (n > 0 && !ic) return fail(CMT_EINVAL, "ic is NULL");
    if (ic_ld < n) return fail(CMT_EINVAL, "ic_ld < n");
    return propagate(bl, false, nullptr, 0, ic, ic_ld, n, first_index, out, workspace, workspace_bytes, stream);
}

extern "C" int cmt_propagate_philox(const cmt_beamline_t *bl, const cmt_source_t *src, uint64_t seed,
                                    int64_t first_index, int64_t n, const cmt_outputs_t *out,
                                    void *workspace, size_t workspace_bytes, void *stream)
{
    if (!src) return fail(CMT_EINVAL, "source is NULL");
    if (src->pos_kind != CMT_POS_DISC && src->pos_kind != CMT_POS_GAUSS)
        return fail(CMT_EINVAL, "unknown pos_kind %d", src->pos_kind);
    return propagate(bl, true, src, seed, nullptr, 0, n, first_index, out, workspace, workspace_bytes, stream);
}

extern "C" int cmt_philox_draw(const cmt_source_t *src, uint64_t seed, int64_t first_index,
                               const int64_t *index, int64_t n, double *ic, int64_t ic_ld, void *stream)
{
    if (!src || (n > 0 && !ic)) return fail(CMT_EINVAL, "NULL argument");
    if (n < 0 || ic_ld < n) return fail(CMT_EINVAL, "bad n / ic_ld");
    if (src->pos_kind != CMT_POS_DISC && src->pos_kind != CMT_POS_GAUSS)
        return fail(CMT_EINVAL, "unknown pos_kind %d", src->pos_kind);
    if (n == 0) return CMT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, n_sm = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)n_sm * 16);
    {
        ScopedTimer tm(3, st);
        draw_kernel<<<grid, 256, 0, st>>>(*src, seed, first_index, index, n, ic, ic_ld);
        count_launch();
    }
    CUDA_TRY(cudaGetLastError());
    return CMT_OK;
}

extern "C" int cmt_trajectories(const cmt_beamline_t *bl, int64_t n, const double *state, int n_comp,
                                int64_t state_ld, const int64_t *select, int64_t select_base, double *rows,
                                int32_t max_rows, const int64_t *row_offset, int32_t *n_rows, uint8_t *fate,
                                void *stream)
{
    if (!bl) return fail(CMT_EINVAL, "beamline handle is NULL");
    if (n < 0) return fail(CMT_EINVAL, "n < 0");
    if (n == 0) return CMT_OK;
    if (!state) return fail(CMT_EINVAL, "state is NULL");
    if (!rows && !n_rows && !fate) return fail(CMT_EINVAL, "nothing to compute: rows, n_rows and fate are all NULL");
    if (n_comp != 6 && n_comp != 10) return fail(CMT_EINVAL, "n_comp must be 6 or 10");
    if (max_rows < 1) return fail(CMT_EINVAL, "max_rows < 1");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(bl->device);
    CUDA_TRY(guard.status);
    const int grid = (int)((n + TRAJ_THREADS - 1) / TRAJ_THREADS);
    {
        ScopedTimer tm(2, st);
        if (bl->math == CMT_MATH_CONTRACTED)
            trajectory_kernel<true><<<grid, TRAJ_THREADS, bl->tab_bytes, st>>>(bl->P, n, state, n_comp, state_ld, select,
                                                                               select_base, rows, max_rows, row_offset, n_rows, fate,
                                                                               nullptr, 0);
        else
            trajectory_kernel<false><<<grid, TRAJ_THREADS, bl->tab_bytes, st>>>(bl->P, n, state, n_comp, state_ld, select,
                                                                                select_base, rows, max_rows, row_offset, n_rows, fate,
                                                                                nullptr, 0);
    }
    CUDA_TRY(cudaGetLastError());
    return CMT_OK;
}

extern "C" int cmt_resume(const cmt_beamline_t *bl, int64_t n, const double *state, int64_t state_ld,
                          double *last_row, int64_t last_ld, int32_t *n_rows, uint8_t *fate, void *stream)
{
    if (!bl) return fail(CMT_EINVAL, "beamline handle is NULL");
    if (n < 0) return fail(CMT_EINVAL, "n < 0");
    if (n == 0) return CMT_OK;
    if (!state) return fail(CMT_EINVAL, "state is NULL");
    if (!last_row && !n_rows && !fate) return fail(CMT_EINVAL, "nothing to compute: last_row, n_rows and fate are all NULL");
    if (state_ld < n || (last_row && last_ld < n)) return fail(CMT_EINVAL, "leading dimension smaller than n");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(bl->device);
    CUDA_TRY(guard.status);
    const int grid = (int)((n + TRAJ_THREADS - 1) / TRAJ_THREADS);
    {
        ScopedTimer tm(2, st);
        if (bl->math == CMT_MATH_CONTRACTED)
            trajectory_kernel<true><<<grid, TRAJ_THREADS, bl->tab_bytes, st>>>(bl->P, n, state, 10, state_ld, nullptr, 0, nullptr, 1,
                                                                               nullptr, n_rows, fate, last_row, last_ld);
        else
            trajectory_kernel<false><<<grid, TRAJ_THREADS, bl->tab_bytes, st>>>(bl->P, n, state, 10, state_ld, nullptr, 0, nullptr, 1,
                                                                                nullptr, n_rows, fate, last_row, last_ld);
        count_launch();
    }
    CUDA_TRY(cudaGetLastError());
    return CMT_OK;
}

extern "C" int cmt_plane_crossings(const cmt_beamline_t *bl, int64_t n, const double *state, int n_comp,
                                   int64_t state_ld, const int64_t *select, int64_t select_base,
                                   const double *z_planes, int32_t n_planes, double *out, int64_t out_ld,
                                   uint8_t *valid, uint8_t *fate, void *stream)
{
    if (!bl) return fail(CMT_EINVAL, "beamline handle is NULL");
    if (n < 0) return fail(CMT_EINVAL, "n < 0");
    if (n_planes < 1 || n_planes > CMT_MAX_PLANES) return fail(CMT_EINVAL, "n_planes must be 1..%d", CMT_MAX_PLANES);
    if (!z_planes) return fail(CMT_EINVAL, "z_planes is NULL");
    ProbePlanes planes;
    memset(&planes, 0, sizeof(planes));
    planes.n = n_planes;
    for (int q = 0; q < n_planes; ++q) {
        if (z_planes[q] != z_planes[q]) return fail(CMT_EINVAL, "z_planes[%d] is NaN", q);
        if (q > 0 && z_planes[q] < z_planes[q - 1]) return fail(CMT_EINVAL, "z_planes must be ascending");
        planes.z[q] = z_planes[q];
    }
    if (n == 0) return CMT_OK;
    if (!state || !out || !valid) return fail(CMT_EINVAL, "state, out and valid are required");
    if (n_comp != 6 && n_comp != 10) return fail(CMT_EINVAL, "n_comp must be 6 or 10");
    if (out_ld < n) return fail(CMT_EINVAL, "out_ld < n");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceGuard guard(bl->device);
    CUDA_TRY(guard.status);
    const int grid = (int)((n + TRAJ_THREADS - 1) / TRAJ_THREADS);
    {
        ScopedTimer tm(2, st);
        if (bl->math == CMT_MATH_CONTRACTED)
            crossing_kernel<true><<<grid, TRAJ_THREADS, bl->tab_bytes, st>>>(bl->P, planes, n, state, n_comp, state_ld, select,
                                                                             select_base, out, out_ld, valid, fate);
        else
            crossing_kernel<false><<<grid, TRAJ_THREADS, bl->tab_bytes, st>>>(bl->P, planes, n, state, n_comp, state_ld, select,
                                                                              select_base, out, out_ld, valid, fate);
    }
    CUDA_TRY(cudaGetLastError());
    return CMT_OK;
}

// ---------------------------------------------------------------------------
// host-buffer entry points: chunked, two streams, copies straight from/to the
// caller's buffers (truly asynchronous when those are pinned, still correct
// when they are pageable).  H2D of chunk k+1 overlaps the kernels of chunk k.
// ---------------------------------------------------------------------------
constexpr int PIPE_SLOTS = 2;   // streams (each with its own buffers and queue workspace) consecutive chunks rotate over; measured
                                // with 4 (round 2): host-IC path 1.106e9 against 1.102e9 molecules/s (the PCIe link is the bound),
                                // a lone Philox run in 3 / 4 pieces on 3 / 4 streams 0.626 / 0.636 ms against 0.583 ms in two
struct HostPipe {
    int device = -1;
    int64_t chunk = 0;
    cudaStream_t st[PIPE_SLOTS] = {};
    double *d_ic[PIPE_SLOTS] = {};
    uint8_t *d_fate[PIPE_SLOTS] = {};
    double *d_final[PIPE_SLOTS] = {};
    void *d_ws[PIPE_SLOTS] = {};
    size_t ws_bytes = 0;
    int64_t *d_cnt = nullptr;   // [CMT_MAX_FATES + CMT_WORK_SLOTS]
    int64_t *h_cnt = nullptr;   // the same, page-locked: the one read-back of a run
    cudaEvent_t ev[PIPE_SLOTS] = {};   // ordering between the slot streams without a host round trip

    void release()
    {
        if (device < 0) return;
        DeviceGuard guard(device);
        for (int k = 0; k < PIPE_SLOTS; ++k) {
            if (st[k]) cudaStreamDestroy(st[k]);
            if (ev[k]) cudaEventDestroy(ev[k]);
            ev[k] = nullptr;
            cudaFree(d_ic[k]); cudaFree(d_fate[k]); cudaFree(d_final[k]); cudaFree(d_ws[k]);
            st[k] = nullptr; d_ic[k] = nullptr; d_fate[k] = nullptr; d_final[k] = nullptr; d_ws[k] = nullptr;
        }
        cudaFree(d_cnt);
        d_cnt = nullptr;
        if (h_cnt) cudaFreeHost(h_cnt);
        h_cnt = nullptr;
        device = -1;
    }
    ~HostPipe() { release(); }
};

static void pipe_destroy(HostPipe *p) { delete p; }

namespace {

int pipe_allocate(HostPipe &p, int64_t chunk, bool keep_ic, bool keep_fate, bool keep_final)
{
    for (int k = 0; k < PIPE_SLOTS; ++k) {
        CUDA_TRY(cudaStreamCreateWithFlags(&p.st[k], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&p.ev[k], cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc(&p.d_ws[k], p.ws_bytes));
        if (keep_ic) CUDA_TRY(cudaMalloc(&p.d_ic[k], (size_t)6 * chunk * sizeof(double)));
        if (keep_fate) CUDA_TRY(cudaMalloc(&p.d_fate[k], (size_t)chunk));
        if (keep_final) CUDA_TRY(cudaMalloc(&p.d_final[k], (size_t)10 * chunk * sizeof(double)));
    }
    CUDA_TRY(cudaMalloc(&p.d_cnt, (CMT_MAX_FATES + CMT_WORK_SLOTS) * sizeof(int64_t)));
    CUDA_TRY(cudaHostAlloc(&p.h_cnt, (CMT_MAX_FATES + CMT_WORK_SLOTS) * sizeof(int64_t), cudaHostAllocDefault));
    return CMT_OK;
}

int pipe_prepare(HostPipe &p, const cmt_beamline_t *bl, int64_t chunk, bool want_ic, bool want_fate, bool want_final)
{
    const size_t ws = cmt_workspace_bytes(bl, chunk);
    const bool ok = p.device == bl->device && p.chunk >= chunk && p.ws_bytes >= cmt_workspace_bytes(bl, p.chunk) &&
                    (!want_final || p.d_final[0]) && (!want_ic || p.d_ic[0]) && (!want_fate || p.d_fate[0]);
    if (ok) return CMT_OK;
    const bool keep_ic = want_ic || p.d_ic[0], keep_fate = want_fate || p.d_fate[0], keep_final = want_final || p.d_final[0];
    chunk = std::max(chunk, p.device == bl->device ? p.chunk : (int64_t)0);
    p.release();
    p.device = bl->device;
    p.chunk = chunk;
    p.ws_bytes = std::max(ws, cmt_workspace_bytes(bl, chunk));
    const int rc = pipe_allocate(p, chunk, keep_ic, keep_fate, keep_final);
    if (rc != CMT_OK) {
        // nothing half-allocated survives a failure: the next call starts from scratch
        p.release();
        p.chunk = 0;
        p.ws_bytes = 0;
    }
    return rc;
}

// Start of a run: the run's Counter is cleared on slot 0 and the other slots wait for that on the device.
int pipe_begin(HostPipe &p)
{
    CUDA_TRY(cudaMemsetAsync(p.d_cnt, 0, (CMT_MAX_FATES + CMT_WORK_SLOTS) * sizeof(int64_t), p.st[0]));
    CUDA_TRY(cudaEventRecord(p.ev[0], p.st[0]));
    for (int k = 1; k < PIPE_SLOTS; ++k) CUDA_TRY(cudaStreamWaitEvent(p.st[k], p.ev[0], 0));
    return CMT_OK;
}

// End of a run: slot 0 waits for the other slots on the device, copies the Counter into page-locked memory, and the
// host waits once.  (Buffers the caller handed in are complete then as well: every copy into them was queued on
// one of the slots.)
int pipe_collect(HostPipe &p, const cmt_beamline_t *bl, int64_t *counters_host, int64_t *work_host)
{
    for (int k = 1; k < PIPE_SLOTS; ++k) {
        CUDA_TRY(cudaEventRecord(p.ev[k], p.st[k]));
        CUDA_TRY(cudaStreamWaitEvent(p.st[0], p.ev[k], 0));
    }
    CUDA_TRY(cudaMemcpyAsync(p.h_cnt, p.d_cnt, (CMT_MAX_FATES + CMT_WORK_SLOTS) * sizeof(int64_t), cudaMemcpyDeviceToHost, p.st[0]));
    CUDA_TRY(cudaStreamSynchronize(p.st[0]));
    const int64_t *h = p.h_cnt;
    for (int f = 0; f < bl->P.n_fates; ++f) counters_host[f] += h[f];
    if (work_host) for (int k = 0; k < CMT_WORK_SLOTS; ++k) work_host[k] += h[CMT_MAX_FATES + k];
    return CMT_OK;
}

}  // namespace

extern "C" int cmt_run_host_ic(const cmt_beamline_t *bl, int64_t n, const double *ic_host, uint8_t *fate_host,
                               double *final_host, int64_t *counters_host, int64_t *work_host)
{
    if (!bl) return fail(CMT_EINVAL, "beamline handle is NULL");
    if (n < 0) return fail(CMT_EINVAL, "n < 0");
    if (!counters_host) return fail(CMT_EINVAL, "counters_host is NULL");
    if (n == 0) return CMT_OK;
    if (!ic_host) return fail(CMT_EINVAL, "ic_host is NULL");
    cmt_beamline_t *owner = const_cast<cmt_beamline_t *>(bl);          // the pipe is a cache, not part of the beamline's value
    std::lock_guard<std::mutex> lk(owner->pipe_mu);
    DeviceGuard guard(bl->device);
    CUDA_TRY(guard.status);
    if (!owner->pipe) owner->pipe = new HostPipe();
    HostPipe &p = *owner->pipe;
    const int64_t chunk = std::min<int64_t>(n, (int64_t)1 << 21);
    int rc = pipe_prepare(p, bl, chunk, true, fate_host != nullptr, final_host != nullptr);
    if (rc) return rc;
    rc = pipe_begin(p);
    if (rc) return rc;

    const size_t dpitch = (size_t)p.chunk * sizeof(double), hpitch = (size_t)n * sizeof(double);
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    static const int ic_slots = std::max(1, std::min(PIPE_SLOTS, env_int("CMT_TUNE_IC_SLOTS", PIPE_SLOTS)));   // experiments only
    for (int64_t ci = 0; ci < n_chunks; ++ci) {
        const int k = (int)(ci % ic_slots);
        const int64_t off = ci * chunk, len = std::min<int64_t>(chunk, n - off);
        CUDA_TRY(cudaMemcpy2DAsync(p.d_ic[k], dpitch, ic_host + off, hpitch, (size_t)len * sizeof(double), 6,
                                   cudaMemcpyHostToDevice, p.st[k]));
        cmt_outputs_t O;
        memset(&O, 0, sizeof(O));
        O.fate = fate_host ? p.d_fate[k] : nullptr;
        O.final_state = final_host ? p.d_final[k] : nullptr;
        O.final_ld = p.chunk;
        O.counters = p.d_cnt;
        O.work = p.d_cnt + CMT_MAX_FATES;
        rc = cmt_propagate_ic(bl, len, off, p.d_ic[k], p.chunk, &O, p.d_ws[k], p.ws_bytes, p.st[k]);
        if (rc) return rc;
        if (fate_host)
            CUDA_TRY(cudaMemcpyAsync(fate_host + off, p.d_fate[k], (size_t)len, cudaMemcpyDeviceToHost, p.st[k]));
        if (final_host)
            CUDA_TRY(cudaMemcpy2DAsync(final_host + off, hpitch, p.d_final[k], dpitch, (size_t)len * sizeof(double),
                                       10, cudaMemcpyDeviceToHost, p.st[k]));
    }
    return pipe_collect(p, bl, counters_host, work_host);
}

extern "C" int cmt_run_host_philox(const cmt_beamline_t *bl, const cmt_source_t *src, uint64_t seed,
                                   int64_t first_index, int64_t n, int64_t *counters_host, int64_t *work_host)
{
    if (!bl) return fail(CMT_EINVAL, "beamline handle is NULL");
    if (n < 0) return fail(CMT_EINVAL, "n < 0");
    if (!counters_host || !src) return fail(CMT_EINVAL, "NULL argument");
    if (n == 0) return CMT_OK;
    cmt_beamline_t *owner = const_cast<cmt_beamline_t *>(bl);
    std::lock_guard<std::mutex> lk(owner->pipe_mu);
    DeviceGuard guard(bl->device);
    CUDA_TRY(guard.status);
    if (!owner->pipe) owner->pipe = new HostPipe();
    HostPipe &p = *owner->pipe;
    // One chunk per 2^26 molecules (8.6 GB of queue workspace per stream), consecutive chunks on
    // alternating streams.  A run that fits one chunk is cut in two halves, one per stream, so that the
    // walk kernel of the second half overlaps the lens segments of the first (1e7 molecules: 1.32e10 ->
    // 1.40e10 molecules/s); three or more pieces were measured slower (the segments are latency-bound,
    // 128 us each whatever their load).  Results do not depend on the cut: the source is indexed by the
    // global molecule number.
    int64_t chunk = std::min<int64_t>(n, (int64_t)1 << 26);
    static const int tune_pieces = env_int("CMT_TUNE_PHILOX_PIECES", 2);          // experiments only
    if (n <= chunk && n >= ((int64_t)1 << 21)) chunk = (n + tune_pieces - 1) / std::max(1, tune_pieces);
    if (const char *env = getenv("CMT_PHILOX_CHUNK")) {          // experiments only
        const long long v = atoll(env);
        if (v > 0) chunk = std::min<int64_t>(n, (int64_t)v);
    }
    int rc = pipe_prepare(p, bl, chunk, false, false, false);
    if (rc) return rc;
    rc = pipe_begin(p);
    if (rc) return rc;
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    for (int64_t ci = 0; ci < n_chunks; ++ci) {
        const int k = (int)(ci % PIPE_SLOTS);
        const int64_t off = ci * chunk, len = std::min<int64_t>(chunk, n - off);
        cmt_outputs_t O;
        memset(&O, 0, sizeof(O));
        O.counters = p.d_cnt;
        O.work = p.d_cnt + CMT_MAX_FATES;
        rc = cmt_propagate_philox(bl, src, seed, first_index + off, len, &O, p.d_ws[k], p.ws_bytes, p.st[k]);
        if (rc) return rc;
    }
    return pipe_collect(p, bl, counters_host, work_host);
}

// ---------------------------------------------------------------------------
// arithmetic self-test
// ---------------------------------------------------------------------------
extern "C" int cmt_selftest(int device, int64_t n, uint64_t seed, int mode, int64_t out[5])
{
    if (!out || n < 0 || mode < 0 || mode > 3) return fail(CMT_EINVAL, "bad selftest arguments");
    DeviceGuard guard(device);
    CUDA_TRY(guard.status);
    unsigned long long *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 5 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(d, 0, 5 * sizeof(unsigned long long)));
    int n_sm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    selftest_kernel<<<n_sm * 8, 256>>>(n, seed, mode, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    unsigned long long h[5] = {0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(CMT_ECUDA, "selftest failed: %s", cudaGetErrorString(e));
    for (int k = 0; k < 5; ++k) out[k] = (int64_t)h[k];
    return CMT_OK;
}

// ---------------------------------------------------------------------------
// FP64 pipe ceilings
// ---------------------------------------------------------------------------
extern "C" int cmt_fp64_peak(int device, double *dfma_per_s, double *dadd_per_s)
{
    DeviceGuard guard(device);
    CUDA_TRY(guard.status);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    double *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(double)));
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    const int grid = prop.multiProcessorCount * 8, iters = 1 << 15;
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 4; ++rep) {
            CUDA_TRY(cudaEventRecord(a));
            if (which == 0) fp64_peak_kernel<true><<<grid, 256>>>(d_out, iters, 1.0);
            else fp64_peak_kernel<false><<<grid, 256>>>(d_out, iters, 1.0);
            CUDA_TRY(cudaEventRecord(b));
            CUDA_TRY(cudaEventSynchronize(b));
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
            const double ops = (double)grid * 256.0 * iters * 8.0;
            if (rep > 0) best[which] = std::max(best[which], ops / (ms * 1e-3));
        }
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    if (dfma_per_s) *dfma_per_s = best[0];
    if (dadd_per_s) *dadd_per_s = best[1];
    return CMT_OK;
}
