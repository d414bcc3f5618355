// Planner, plan cache and pass sequencing (host C++).
//
// The reference creates and destroys two cuFFT plans and four device buffers inside every call
// (/root/reference/src/convolution3Dfft.cu:442-559).  Here a plan (twiddle / permutation tables,
// spectrum workspace, stream) is built once per (device, shape) and cached, thread-safe.
#include "fc_plan.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace fcb200 {

// ------------------------------------------------------------------------------------------------
// pure host planning
// ------------------------------------------------------------------------------------------------
// FCB200_PLAN="z300=20.15,x280=20.14,...": radix sequence of one axis style (y, z, x) and length, for tuning runs.
// Lengths without a compile-time kernel for the given sequence run on the run-time-radix kernels.
static bool plan_override(int L, int style, std::vector<int>& out, bool& gen)
{
    const char* e = std::getenv("FCB200_PLAN");
    if (!e) return false;
    const char tag = style == 1 ? 'z' : (style == 2 ? 'x' : 'y');
    std::string all(e);
    size_t at = 0;
    while (at < all.size()) {
        size_t end = all.find(',', at);
        if (end == std::string::npos) end = all.size();
        const std::string item = all.substr(at, end - at);
        at = end + 1;
        const size_t eq = item.find('=');
        if (item.size() < 3 || item[0] != tag || eq == std::string::npos || std::atoi(item.c_str() + 1) != L) continue;
        std::vector<int> rad;
        long long prod = 1;
        size_t q = eq + 1;
        while (q < item.size()) {
            const int r = std::atoi(item.c_str() + q);
            if (r < 2) break;
            rad.push_back(r);
            prod *= r;
            q = item.find('.', q);
            if (q == std::string::npos) break;
            ++q;
        }
        if (prod != L || rad.empty() || (int)rad.size() > kMaxStages) continue;
        static const int fast[] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 11, 13, 17, 19, 23, 14, 18, 20, 21, 25, 28};
        gen = false;
        for (int r : rad)
            if (std::find(std::begin(fast), std::end(fast), r) == std::end(fast)) gen = true;
        out = rad;
        return true;
    }
    return false;
}

std::vector<int> factorize(int L, bool* generic, int style)
{
    std::vector<int> out;
    bool gen = false;
    if (plan_override(L, style, out, gen)) {
        if (generic) *generic = gen;
        return out;
    }
    // Two stages of FAT composite radices (fft_butterflies.cuh: Dft<14> .. Dft<28>, Cooley-Tukey inside the registers of one
    // thread) for the extents of the caller-padded configurations whose prime-by-prime plans need three or four shared-memory
    // round trips -- each measured against the plan it replaces (profiles/r02_plan_ab.jsonl; y passes / fused z pass):
    //   L = 300 (4,3,5,5) -> (20,15): 0.101 -> 0.065 / 0.289 -> 0.215 ms    L = 420 (4,3,5,7) -> (20,21): 0.153 -> 0.102 / 0.318 -> 0.181 ms
    //   L = 270 (2,15,9)  -> (18,15): 0.041 -> 0.027 / 0.064 -> 0.056 ms    L = 448 y only (8,8,7) -> (16,28): 0.109 -> 0.085 ms
    //   L = 560 fused z only (16,5,7) -> (28,20): 0.205 -> 0.177 ms          x half-lengths 150 -> (10,15), 135 -> (9,15): -3..8 %
    // (a radix-28 butterfly around the spectrum multiply loses: z 448 stays (8,8,7); the tiled x kernels lose with
    // (20,14) / (16,14) / (14,15): 280, 224, 210 keep their primes.)
    struct FatPlan { int L, styles, r0, r1; };   // styles: bit 0 y axis, bit 1 fused z axis, bit 2 x axis
    static const FatPlan fat[] = {{300, 3, 20, 15}, {420, 3, 20, 21}, {270, 3, 18, 15}, {448, 1, 16, 28},
                                  {560, 2, 28, 20}, {150, 4, 10, 15}, {135, 4, 9, 15}};
    for (const FatPlan& f : fat)
        if (f.L == L && ((f.styles >> style) & 1)) {
            if (generic) *generic = false;
            return {f.r0, f.r1};
        }
    int n = L;
    // power-of-two part: fewest stages with radices <= 16, never a trailing radix 2 when avoidable
    int e = 0;
    while (n > 1 && (n & 1) == 0) {
        n >>= 1;
        ++e;
    }
    // style 0 (y axis): radix 16 where it saves a stage.  style 2 (x axis): the same except L = 1024.
    // style 1 (fused z axis): the measured
    // exception L = 256 -> (8,8,4), because the fused forward-multiply-inverse kernel is register-bound
    // with two radix-16 stages (profiles/r01_notes.md).
    static const int pow2_plan[13][4] = {
        {0, 0, 0, 0},    {2, 0, 0, 0},    {4, 0, 0, 0},    {8, 0, 0, 0},   {16, 0, 0, 0},  {8, 4, 0, 0},   {8, 8, 0, 0},
        {16, 8, 0, 0},   {16, 16, 0, 0},  {8, 8, 8, 0},    {16, 16, 4, 0}, {16, 16, 8, 0}, {16, 16, 16, 0}};
    while (e > 12) {
        out.push_back(16);
        e -= 4;
    }
    if (style == 1 && L == 256) {
        out = {8, 8, 4};
    } else if (style == 2 && L == 1024) {
        out = {16, 8, 8};   // x axis: the row-wise register kernel needs stage strides that are multiples of 8
    } else if (style == 2 && L == 512) {
        out = {16, 4, 8};   // x axis (nx = 1024): 32 lanes per row pair, warp-level synchronisation only
    } else {
        for (int i = 0; i < 4 && pow2_plan[e][i]; ++i) out.push_back(pow2_plan[e][i]);
    }
    // odd 7-smooth part.  Up to four stages in total: one register stage per prime (3, 5, 7).  Lengths that would
    // need FIVE or more shared-memory round trips (270 = 2*3*3*3*5, 1125, 2160 ...) pair their small primes into
    // composite register stages instead (radix 15 = 3*5, 9 = 3*3; a single leftover 3 or 5 joins a trailing radix
    // 2 / 4 -> 6, 12, 10): 270 = (2,15,9), 1125 = (15,15,5), 2160 = (16,15,9).  Measured (profiles/r01_notes.md):
    // 270^3 0.438 -> 0.390 ms, whereas the four-stage lengths 300 / 420 lose with a radix-15 stage and keep primes.
    int n3 = 0, n5 = 0, n7 = 0;
    while (n % 3 == 0) { n /= 3; ++n3; }
    while (n % 5 == 0) { n /= 5; ++n5; }
    while (n % 7 == 0) { n /= 7; ++n7; }
    // (x axis: the tiled kernels keep one stage per prime up to four stages, they lose with composite radices; y / z
    // axes: at most three stages where pairing allows it -- 360 = (8,15,3), 480 = (8,4,15), 288 = (8,4,9), 350 = (10,5,7) --
    // which is also what lets their digit reversal fit the five dimensions of a tensor map, fft_col_tma.cu)
    if ((int)out.size() + n3 + n5 + n7 > (style == 2 ? 4 : 3)) {
        std::vector<int> comp;
        while (n3 >= 1 && n5 >= 1) { comp.push_back(15); --n3; --n5; }
        if ((n3 & 1) && !out.empty() && (out.back() == 2 || out.back() == 4)) { out.back() *= 3; --n3; }
        while (n3 >= 2) { comp.push_back(9); n3 -= 2; }
        if (n5 >= 1 && !out.empty() && out.back() == 2) { out.back() = 10; --n5; }
        out.insert(out.end(), comp.begin(), comp.end());
    }
    out.insert(out.end(), (size_t)n3, 3);
    out.insert(out.end(), (size_t)n5, 5);
    out.insert(out.end(), (size_t)n7, 7);
    for (int p = 11; n > 1; p += 2) {
        if ((long long)p * p > n) p = n;  // remaining cofactor is prime
        while (n % p == 0) {
            out.push_back(p);
            n /= p;
            if (p > 23) gen = true;   // 11, 13, 17, 19, 23 have register butterflies (fft_butterflies.cuh: DftOddPrime)
        }
    }
    if (out.empty()) out.push_back(1);
    if ((int)out.size() > kMaxStages) throw std::runtime_error("fcb200: transform length has too many factors");
    if (generic) *generic = gen;
    return out;
}

void build_tables(int L, const std::vector<int>& radix, std::vector<int>& rev, std::vector<int>& pos,
                  std::vector<float2>& tw)
{
    rev.assign(L, 0);
    pos.assign(L, 0);
    tw.resize(L);
    for (int p = 0; p < L; ++p) {
        int k = 0, mul = 1, rem = p, Li = L;
        for (int R : radix) {
            int S = Li / R;
            int m = rem / S;
            rem -= m * S;
            k += m * mul;
            mul *= R;
            Li = S;
        }
        rev[p] = k;
        pos[k] = p;
    }
    const double two_pi = 6.283185307179586476925286766559;
    for (int t = 0; t < L; ++t) {
        // exact symmetries first so the eight principal roots are exact in float
        double ang = two_pi * (double)t / (double)L;
        double c = std::cos(ang), s = -std::sin(ang);
        if (4 * (long long)t == L) { c = 0.0; s = -1.0; }
        if (2 * (long long)t == L) { c = -1.0; s = 0.0; }
        if (4 * (long long)t == 3LL * L) { c = 0.0; s = 1.0; }
        tw[t] = make_float2((float)c, (float)s);
    }
}

Geometry make_geometry(int nx, int ny, int nz)
{
    Geometry g{};
    g.nx = nx;
    g.ny = ny;
    g.nz = nz;
    g.xc = nx / 2 + 1;
    g.xcp = (g.xc + 3) & ~3;
    g.odd = nx & 1;
    g.M = g.odd ? nx : nx / 2;
    return g;
}

std::vector<int> psf_active_rows(const int* dims, const int* kdims, int nx)
{
    const long long d0 = dims[0], d1 = dims[1], d2 = dims[2];
    const int k0 = kdims[0], k1 = kdims[1], k2 = kdims[2];
    const long long nrows = d0 * d1 * d2 / nx;
    std::vector<unsigned char> hit((size_t)nrows, 0);
    for (int a = 0; a < k0; ++a) {
        long long aq = a - k0 / 2;
        if (aq < 0) aq += d0;
        for (int b = 0; b < k1; ++b) {
            long long bq = b - k1 / 2;
            if (bq < 0) bq += d1;
            for (int c = 0; c < k2; ++c) {
                long long cq = c - k2 / 2;
                if (cq < 0) cq += d2;
                long long flat = cq + d2 * (bq + d1 * aq);
                hit[(size_t)(flat / nx)] = 1;
            }
        }
    }
    std::vector<int> rows;
    for (long long r = 0; r < nrows; ++r)
        if (hit[(size_t)r]) rows.push_back((int)r);
    return rows;
}

// ------------------------------------------------------------------------------------------------
// device plan
// ------------------------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }
void count_launches(int n) { g_launches.fetch_add(n); }

// PDL pays in the launch-bound regime (64^3: -6 %, 256^3: -3 %); on large volumes the early-resident CTAs of the next
// pass take SMs away from the PSF passes running on the side stream (C3: +6 %, profiles/r01_sweep_pdl.jsonl), so
// it is used for plans whose spectrum fits the L2 only
static int plan_pdl(const ConvPlan& p) { return (pdl_enabled() && p.spec_bytes() <= (size_t)96 << 20) ? 1 : 0; }

bool pdl_enabled()
{
    static const bool on = [] {
        const char* e = std::getenv("FCB200_PDL");
        return !(e && std::atoi(e) == 0);
    }();
    return on;
}

// ---- optional per-pass timing with CUDA events on the launching stream (bench.py roofline) ----
static std::atomic<int> g_profile{0};
struct PassEvent {
    int id;
    cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static std::vector<PassEvent> g_prof_events;

void profile_enable(int on) { g_profile.store(on ? 1 : 0); }
bool profile_enabled() { return g_profile.load() != 0; }

int profile_read(float* ms_sum, long long* counts, int n)
{
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (int i = 0; i < n; ++i) {
        ms_sum[i] = 0.f;
        counts[i] = 0;
    }
    for (PassEvent& e : g_prof_events) {
        float ms = 0.f;
        if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess &&
            e.id >= 0 && e.id < n) {
            ms_sum[e.id] += ms;
            counts[e.id] += 1;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof_events.clear();
    return kNumPassIds;
}

struct PassTimer {
    int id;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    PassTimer(int id_, cudaStream_t st_) : id(id_), st(st_)
    {
        if (g_profile.load()) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
    }
    void cancel()   // nothing was launched: drop the events
    {
        if (a) {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
            a = b = nullptr;
        }
    }
    ~PassTimer()
    {
        if (a) {
            cudaEventRecord(b, st);
            std::lock_guard<std::mutex> lock(g_prof_mu);
            g_prof_events.push_back({id, a, b});
        }
    }
};

bool build_rader(int p, RaderTables& r)
{
    r = RaderTables();
    if (p < 3) return false;
    const int n = p - 1;
    bool gen = false;
    std::vector<int> radix = factorize(n, &gen, 3);   // style 3: y-axis planning without the two-stage fat plans
    for (int R : radix)
        if (!(R == 2 || R == 3 || R == 4 || R == 5 || R == 6 || R == 7 || R == 8 || R == 9 || R == 10 || R == 11 || R == 12 ||
              R == 13 || R == 15 || R == 16))
            return false;
    if (gen || radix.size() > 8) return false;
    // smallest primitive root g of p
    auto powmod = [&](long long b, long long e) {
        long long x = 1 % p;
        b %= p;
        while (e > 0) {
            if (e & 1) x = x * b % p;
            b = b * b % p;
            e >>= 1;
        }
        return x;
    };
    std::vector<int> prime_factors;
    {
        int m = n;
        for (int q = 2; (long long)q * q <= m; ++q)
            if (m % q == 0) {
                prime_factors.push_back(q);
                while (m % q == 0) m /= q;
            }
        if (m > 1) prime_factors.push_back(m);
    }
    int g = 0;
    for (int c = 2; c < p && !g; ++c) {
        bool ok = true;
        for (int q : prime_factors) ok = ok && powmod(c, n / q) != 1;
        if (ok) g = c;
    }
    if (!g) return false;
    r.p = p;
    r.n = n;
    r.radix = radix;
    r.perm.resize(n);
    r.iperm.resize(n);
    const long long ginv = powmod(g, p - 2);
    long long x = 1, y = 1;
    for (int m = 0; m < n; ++m) {
        r.perm[m] = (int)x;
        r.iperm[m] = (int)y;
        x = x * g % p;
        y = y * ginv % p;
    }
    std::vector<int> rev, pos;
    build_tables(n, radix, rev, pos, r.tw);
    // spectra of b[t] = exp(-+2*pi*i*g^(-t)/p), divided by n, stored by POSITION of the n-point DIF (position q holds
    // frequency rev[q]); direct O(n^2) sums in double (n < 5000)
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<double> br(n), bim(n);
    for (int t = 0; t < n; ++t) {
        const double ang = two_pi * (double)r.iperm[t] / (double)p;
        br[t] = std::cos(ang);
        bim[t] = -std::sin(ang);   // forward: exp(-i ang)
    }
    r.bf.resize(n);
    r.bi.resize(n);
    for (int q = 0; q < n; ++q) {
        const int f = rev[q];
        double fr = 0, fi = 0, ir = 0, ii = 0;
        for (int t = 0; t < n; ++t) {
            const long long e = ((long long)f * t) % n;
            const double a = two_pi * (double)e / (double)n, c = std::cos(a), sn = -std::sin(a);   // exp(-2 pi i f t / n)
            // forward b = (br, bim); inverse b = conj = (br, -bim)
            fr += br[t] * c - bim[t] * sn;
            fi += br[t] * sn + bim[t] * c;
            ir += br[t] * c + bim[t] * sn;
            ii += br[t] * sn - bim[t] * c;
        }
        r.bf[q] = make_float2((float)(fr / n), (float)(fi / n));
        r.bi[q] = make_float2((float)(ir / n), (float)(ii / n));
    }
    return true;
}

static int rader_min_prime()
{
    static const int v = [] {
        // measured (profiles/r02_odd_sizes.jsonl): 79 / 109 are faster as direct sums (158x158x218: 0.50 against 1.05 ms),
        // 271 is faster with Rader (542x542x296: 9.6 -> 6.7 ms); 0 turns Rader off
        const char* e = std::getenv("FCB200_RADER_MIN");
        return e ? std::atoi(e) : 160;
    }();
    return v;
}

static void make_axis(AxisPlan& a, int L, int style)
{
    a.L = L;
    a.radix = factorize(L, &a.generic, style);
    std::vector<int> rev, pos;
    std::vector<float2> tw;
    build_tables(L, a.radix, rev, pos, tw);
    FC_CUDA(cudaMalloc(&a.d_tw, sizeof(float2) * L));
    FC_CUDA(cudaMalloc(&a.d_rev, sizeof(int) * L));
    FC_CUDA(cudaMalloc(&a.d_pos, sizeof(int) * L));
    FC_CUDA(cudaMemcpy(a.d_tw, tw.data(), sizeof(float2) * L, cudaMemcpyHostToDevice));
    FC_CUDA(cudaMemcpy(a.d_rev, rev.data(), sizeof(int) * L, cudaMemcpyHostToDevice));
    FC_CUDA(cudaMemcpy(a.d_pos, pos.data(), sizeof(int) * L, cudaMemcpyHostToDevice));
    a.dev.L = L;
    a.dev.ns = (int)a.radix.size();
    for (int i = 0; i < kMaxStages; ++i) a.dev.radix[i] = i < a.dev.ns ? a.radix[i] : 1;
    a.dev.generic = a.generic ? 1 : 0;
    a.dev.big = 0;
    for (int r : a.radix)
        if (r == 11 || r == 13 || r == 17 || r == 19 || r == 23 || r == 14 || r == 18 || r == 20 || r == 21 || r == 25 || r == 28)
            a.dev.big = 1;   // butterflies that only the "big" builds of the run-time-radix kernels carry
    a.dev.tw = a.d_tw;
    a.dev.rev = a.d_rev;
    a.dev.pos = a.d_pos;
    // Rader for a large prime as the LAST radix (the planner puts the primes last, ascending)
    a.dev.rader_p = a.dev.rader_n = a.dev.rader_ns = 0;
    const int plast = a.radix.empty() ? 1 : a.radix.back();
    RaderTables rt;
    if (a.generic && plast > 23 && plast >= rader_min_prime() && rader_min_prime() > 0 && build_rader(plast, rt)) {
        const int n = rt.n;
        FC_CUDA(cudaMalloc(&a.d_rtw, sizeof(float2) * n));
        FC_CUDA(cudaMalloc(&a.d_rbf, sizeof(float2) * n));
        FC_CUDA(cudaMalloc(&a.d_rbi, sizeof(float2) * n));
        FC_CUDA(cudaMalloc(&a.d_rperm, sizeof(int) * n));
        FC_CUDA(cudaMalloc(&a.d_riperm, sizeof(int) * n));
        FC_CUDA(cudaMemcpy(a.d_rtw, rt.tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
        FC_CUDA(cudaMemcpy(a.d_rbf, rt.bf.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
        FC_CUDA(cudaMemcpy(a.d_rbi, rt.bi.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
        FC_CUDA(cudaMemcpy(a.d_rperm, rt.perm.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        FC_CUDA(cudaMemcpy(a.d_riperm, rt.iperm.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        a.dev.rader_p = plast;
        a.dev.rader_n = n;
        a.dev.rader_ns = (int)rt.radix.size();
        for (int i = 0; i < 8; ++i) a.dev.rader_radix[i] = i < a.dev.rader_ns ? rt.radix[(size_t)i] : 1;
        a.dev.rader_tw = a.d_rtw;
        a.dev.rader_perm = a.d_rperm;
        a.dev.rader_iperm = a.d_riperm;
        a.dev.rader_bf = a.d_rbf;
        a.dev.rader_bi = a.d_rbi;
        a.dev.big = 1;   // the Rader stage (and the radices 11 / 13 of its n-point transform) live in the "big" kernel builds
    }
}

static void free_axis(AxisPlan& a)
{
    cudaFree(a.d_tw);
    cudaFree(a.d_rev);
    cudaFree(a.d_pos);
    cudaFree(a.d_rtw);
    cudaFree(a.d_rbf);
    cudaFree(a.d_rbi);
    cudaFree(a.d_rperm);
    cudaFree(a.d_riperm);
    a.d_rtw = a.d_rbf = a.d_rbi = nullptr;
    a.d_rperm = a.d_riperm = nullptr;
    a.d_tw = nullptr;
    a.d_rev = a.d_pos = nullptr;
}

ConvPlan::~ConvPlan()
{
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    free_axis(px);
    free_axis(py);
    free_axis(pz);
    for (int i = 0; i < 3; ++i) {
        if (i > 0) cudaFree(d_ring[i]);   // d_ring[0] aliases d_real
        if (ev_up[i]) cudaEventDestroy(ev_up[i]);
        if (ev_comp[i]) cudaEventDestroy(ev_comp[i]);
        if (ev_down[i]) cudaEventDestroy(ev_down[i]);
    }
    for (cudaEvent_t e : ev_chunk)
        if (e) cudaEventDestroy(e);
    if (ev_busy) cudaEventDestroy(ev_busy);
    if (s_psf) cudaStreamDestroy(s_psf);
    if (ev_psf_fork) cudaEventDestroy(ev_psf_fork);
    if (ev_psf_done) cudaEventDestroy(ev_psf_done);
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    cudaFree(d_twx);
    cudaFree(d_spec);
    cudaFree(d_H);
    cudaFree(d_Hwin);
    cudaFree(d_win_slot);
    cudaFree(d_real);
    cudaFree(d_kernel);
    cudaFree(d_unpadded);
    cudaFree(d_rows);
    cudaFree(d_rows_c);
    cudaFree(d_planes);
    cudaFree(d_plane_mask);
    cudaFree(d_tap_start);
    cudaFree(d_tap_x);
    cudaFree(d_tap_idx);
    if (stream) cudaStreamDestroy(stream);
    if (prev >= 0) cudaSetDevice(prev);
}

struct PlanKey {
    int dev, nx, ny, nz, ws;
    bool operator<(const PlanKey& o) const
    {
        if (ws != o.ws) return ws < o.ws;
        if (dev != o.dev) return dev < o.dev;
        if (nx != o.nx) return nx < o.nx;
        if (ny != o.ny) return ny < o.ny;
        return nz < o.nz;
    }
};

static std::mutex g_cache_mu;
static std::map<PlanKey, std::shared_ptr<ConvPlan>> g_cache;
static unsigned long long g_tick = 0;

static size_t max_cached_plans()
{
    const char* e = std::getenv("FCB200_MAX_PLANS");
    if (e) {
        long v = std::strtol(e, nullptr, 10);
        if (v >= 0) return (size_t)v;
    }
    return 4;
}

static void enforce_byte_budget_locked(const ConvPlan* keep);

static void evict_lru_locked(size_t keep)
{
    while (g_cache.size() > keep) {
        auto victim = g_cache.end();
        for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
            if (it->second.use_count() == 1 && (victim == g_cache.end() || it->second->last_use < victim->second->last_use))
                victim = it;
        if (victim == g_cache.end()) break;
        g_cache.erase(victim);
    }
}

static std::shared_ptr<ConvPlan> build_plan(int device, int nx, int ny, int nz, bool workspace)
{
    auto p = std::make_shared<ConvPlan>();
    p->device = device;
    p->g = make_geometry(nx, ny, nz);
    make_axis(p->px, p->g.M, 2);
    make_axis(p->py, ny, 0);
    static const int zstyle = [] {   // experiments: FCB200_ZSTYLE=0 plans the z axis like the y axis (L = 256 as (16,16))
        const char* e = std::getenv("FCB200_ZSTYLE");
        return e ? std::atoi(e) : 1;
    }();
    make_axis(p->pz, nz, zstyle);
    if (!x_pass_supported(p->g, p->px.dev))
        throw std::runtime_error("fcb200: fastest image extent too large for the shared-memory x pass");
    p->txp_y = col_pick_txp(p->py.dev);
    p->txp_z = col_pick_txp(p->pz.dev);
    if (p->txp_y == 0 || p->txp_z == 0)
        throw std::runtime_error("fcb200: image extent too large for the shared-memory column pass");
    if (!p->g.odd) {
        std::vector<float2> twx(p->g.M + 1);
        const double two_pi = 6.283185307179586476925286766559;
        for (int k = 0; k <= p->g.M; ++k) {
            double ang = two_pi * (double)k / (double)nx;
            twx[k] = make_float2((float)std::cos(ang), (float)(-std::sin(ang)));
        }
        FC_CUDA(cudaMalloc(&p->d_twx, sizeof(float2) * twx.size()));
        FC_CUDA(cudaMemcpy(p->d_twx, twx.data(), sizeof(float2) * twx.size(), cudaMemcpyHostToDevice));
    }
    if (workspace) FC_CUDA(cudaMalloc(&p->d_spec, p->spec_bytes()));   // d_H / d_Hwin: on first use
    FC_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    return p;
}

std::shared_ptr<ConvPlan> get_plan(int device, int nx, int ny, int nz, bool workspace)
{
    std::lock_guard<std::mutex> lock(g_cache_mu);
    PlanKey key{device, nx, ny, nz, workspace ? 1 : 0};
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
        it->second->last_use = ++g_tick;
        return it->second;
    }
    std::shared_ptr<ConvPlan> p;
    try {
        p = build_plan(device, nx, ny, nz, workspace);
    } catch (const std::runtime_error&) {
        // most likely out of device memory: drop every idle cached plan and retry once
        cudaGetLastError();
        evict_lru_locked(0);
        p = build_plan(device, nx, ny, nz, workspace);
    }
    p->last_use = ++g_tick;
    g_cache[key] = p;
    size_t keep = max_cached_plans();
    evict_lru_locked(keep == 0 ? 1 : keep);
    enforce_byte_budget_locked(p.get());
    return p;
}

cudaError_t device_alloc_retry(void** p, size_t bytes)
{
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaErrorMemoryAllocation) return e;
    cudaGetLastError();
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        evict_lru_locked(0);   // plans in use are held by their callers (use_count > 1) and survive
    }
    return cudaMalloc(p, bytes);
}

// The cache is bounded by a plan count (FCB200_MAX_PLANS) and by a byte budget (FCB200_CACHE_MB, default: half of the
// device memory): idle least-recently-used plans go first.  The plan just handed out is never evicted.
static size_t plan_bytes(const ConvPlan& p)
{
    size_t b = 0;
    if (p.d_spec) b += p.spec_bytes();
    if (p.d_H) b += p.spec_bytes();
    if (p.d_real) b += p.real_bytes();
    for (int i = 1; i < 3; ++i)
        if (p.d_ring[i]) b += p.real_bytes();
    b += p.unpadded_cap * sizeof(float) + p.hwin_cap * sizeof(float2);
    return b;
}

static void enforce_byte_budget_locked(const ConvPlan* keep)
{
    static const long long budget_mb = [] {
        const char* e = std::getenv("FCB200_CACHE_MB");
        return e ? std::atoll(e) : -1LL;
    }();
    size_t budget;
    if (budget_mb >= 0) {
        budget = (size_t)budget_mb << 20;
    } else {
        // half of the device memory; queried once per device (cudaMemGetInfo costs about a millisecond)
        static size_t total_of[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return;
        if (total_of[dev] == 0) {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
                cudaGetLastError();
                return;
            }
            total_of[dev] = total_b;
        }
        budget = total_of[dev] / 2;
    }
    for (;;) {
        size_t sum = 0;
        for (auto& kv : g_cache) sum += plan_bytes(*kv.second);
        if (sum <= budget) return;
        auto victim = g_cache.end();
        for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
            if (it->second.get() != keep && it->second.use_count() == 1 &&
                (victim == g_cache.end() || it->second->last_use < victim->second->last_use))
                victim = it;
        if (victim == g_cache.end()) return;
        g_cache.erase(victim);
    }
}

void workspace_acquire(ConvPlan& p, cudaStream_t st)
{
    if (p.ev_busy) FC_CUDA(cudaStreamWaitEvent(st, p.ev_busy, 0));
}

void workspace_release(ConvPlan& p, cudaStream_t st)
{
    if (!p.ev_busy) FC_CUDA(cudaEventCreateWithFlags(&p.ev_busy, cudaEventDisableTiming));
    FC_CUDA(cudaEventRecord(p.ev_busy, st));
}

void release_all_plans()
{
    std::lock_guard<std::mutex> lock(g_cache_mu);
    g_cache.clear();
}

// ------------------------------------------------------------------------------------------------
// pass sequencing
// ------------------------------------------------------------------------------------------------
static ColArgs y_args(ConvPlan& p, float2* data)
{
    ColArgs a{};
    a.data = data;
    a.H = nullptr;
    a.P = p.py.dev;
    a.pdl = plan_pdl(p);
    a.stride = p.g.xcp;
    a.groupStride = (long long)p.g.ny * p.g.xcp;
    a.txp = p.txp_y;
    a.tilesPerGroup = (p.g.xcp + 2 * a.txp - 1) / (2 * a.txp);
    a.rowLen = p.g.xcp;
    a.scale = 1.f;
    a.groupList = nullptr;
    a.rowMask = nullptr;
    a.split = nullptr;
    a.splitRows = 0;
    a.splitBlock = a.splitGroup = 0;
    a.winSlot = nullptr;
    a.winPlanes = 0;
    a.splitPeers = nullptr;
    a.splitPeerOffset = 0;
    a.splitInPeers = nullptr;
    return a;
}

static void col_pass(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    if (launch_col_tma(a, mode, ngroups, st)) return;
    if (launch_col_static(a, mode, ngroups, st)) return;
    if (a.split || a.splitPeers) {   // split (exchange-buffer / peer) layout: static kernels or the generic one
        launch_col(a, mode, ngroups, st);
        return;
    }
    if (a.txp == 8 && col_fast_supported(a.P)) launch_col_fast(a, mode, ngroups, st);
    else launch_col(a, mode, ngroups, st);
}

static ColArgs z_args(ConvPlan& p, float2* data)
{
    ColArgs a{};
    const long long C = (long long)p.g.ny * p.g.xcp;
    if (C > 0x7fffffffLL) throw std::runtime_error("fcb200: plane too large");
    a.data = data;
    a.H = nullptr;
    a.P = p.pz.dev;
    a.pdl = plan_pdl(p);
    a.stride = C;
    a.groupStride = 0;
    a.txp = p.txp_z;
    a.tilesPerGroup = (int)((C + 2 * a.txp - 1) / (2 * a.txp));
    a.rowLen = (int)C;
    a.scale = 1.f;
    a.groupList = nullptr;
    a.rowMask = nullptr;
    a.split = nullptr;
    a.splitRows = 0;
    a.splitBlock = a.splitGroup = 0;
    a.winSlot = nullptr;
    a.winPlanes = 0;
    a.splitPeers = nullptr;
    a.splitPeerOffset = 0;
    a.splitInPeers = nullptr;
    return a;
}

static XArgs x_args(ConvPlan& p)
{
    XArgs a{};
    a.g = p.g;
    a.P = p.px.dev;
    a.pdl = plan_pdl(p);
    a.twx = p.d_twx;
    a.nrows = (long long)p.g.ny * p.g.nz;
    a.rowList = nullptr;
    a.compactOut = 0;
    a.txp = 8;
    a.tapStart = a.tapX = a.tapIdx = nullptr;
    return a;
}

void run_forward(ConvPlan& p, const float* d_real, float2* dst, int passes, cudaStream_t st)
{
    XArgs xa = x_args(p);
    xa.in_real = d_real;
    xa.spec = dst;
    {
        PassTimer t(kPassXFwd, st);
        launch_x_fwd(xa, false, st);
    }
    count_launches(1);
    if (passes >= 2) {
        PassTimer t(kPassYFwd, st);
        col_pass(y_args(p, dst), 0, p.g.nz, st);
        count_launches(1);
    }
    if (passes >= 3) {
        PassTimer t(kPassPsfZ, st);
        col_pass(z_args(p, dst), 0, 1, st);
        count_launches(1);
    }
}

// Planes of the padded PSF volume that the PSF passes work on.  mask[z] = 1: plane z holds at least one tap.
// When every such plane lies inside a window of 16, 32 or 64 consecutive planes (mod nz) the list is that window in
// WINDOW order (list entry n = plane (z0 + n) mod nz, tapless planes included: they transform to zeros), so that the
// compact buffers of the on-the-fly path are indexed by window position; otherwise the planes with taps, ascending.
static void psf_plane_list(const int* pdims, const Geometry& g, std::vector<unsigned char>& mask, std::vector<int>& planes,
                           int& win_z0, int& win_planes)
{
    std::vector<int> arows = psf_active_rows(pdims + 3, pdims, g.nx);
    mask.assign((size_t)g.nz, 0);
    for (int r : arows) mask[(size_t)(r / g.ny)] = 1;
    std::vector<int> active;
    for (int z = 0; z < g.nz; ++z)
        if (mask[(size_t)z]) active.push_back(z);
    win_z0 = -1;
    win_planes = 0;
    for (int wp = 16; wp <= 64 && wp <= g.nz && win_z0 < 0; wp *= 2)
        for (int z0 : active) {
            bool ok = true;
            for (int z : active) ok = ok && (((z - z0) % g.nz + g.nz) % g.nz < wp);
            if (ok) {
                win_z0 = z0;
                win_planes = wp;
                break;
            }
        }
    planes.clear();
    if (win_z0 >= 0)
        for (int n = 0; n < win_planes; ++n) planes.push_back((win_z0 + n) % g.nz);
    else
        planes = active;
}

// PSF pruning lists (cached per placement dims): only z planes that receive a tap are non-zero before
// the z pass.  The x pass runs on every row of those planes (rows without taps transform to zeros, so
// nothing needs clearing), the y pass on those planes only, and the z pass reads only those planes.
static void psf_lists(ConvPlan& p, const int* pdims, cudaStream_t st)
{
    if (std::memcmp(p.psf_key, pdims, sizeof(int) * 6) != 0 || p.d_rows == nullptr) {
        std::vector<unsigned char> mask;
        std::vector<int> planes, rows;
        int win_z0 = -1, win_planes = 0;
        psf_plane_list(pdims, p.g, mask, planes, win_z0, win_planes);
        // the x pass transforms only the rows that hold a tap (a few dozen of the ny rows of a plane: C3 31 of 512,
        // C5 63 of 2048); the other rows of the listed planes are zero and are cleared with a memset instead.
        // rows[i]: global row z*ny + y (output row of the materialised path); rows_c[i]: row in a compact buffer that
        // holds the listed planes back to back (window buffer, slab scratch): (list plane)*ny + y
        std::vector<int> rows_c, row_index((size_t)planes.size() * p.g.ny, -1);
        {
            std::vector<int> plane_slot0((size_t)p.g.nz, -1);
            for (size_t i = 0; i < planes.size(); ++i) plane_slot0[(size_t)planes[i]] = (int)i;
            std::vector<unsigned char> has((size_t)planes.size() * p.g.ny, 0);
            for (int r : psf_active_rows(pdims + 3, pdims, p.g.nx)) {
                const int slot = plane_slot0[(size_t)(r / p.g.ny)];
                if (slot >= 0) has[(size_t)slot * p.g.ny + (size_t)(r % p.g.ny)] = 1;
            }
            for (size_t slot = 0; slot < planes.size(); ++slot)
                for (int y = 0; y < p.g.ny; ++y)
                    if (has[slot * p.g.ny + (size_t)y]) {
                        row_index[slot * p.g.ny + (size_t)y] = (int)rows.size();
                        rows.push_back(planes[slot] * p.g.ny + y);
                        rows_c.push_back((int)slot * p.g.ny + y);
                    }
        }
        FC_CUDA(cudaStreamSynchronize(st));  // earlier launches on this stream may still read the old lists
        std::memset(p.psf_key, 0, sizeof(p.psf_key));   // a failed rebuild must not leave a key that matches stale lists
        if (rows.size() > p.rows_cap) {
            cudaFree(p.d_rows);
            cudaFree(p.d_rows_c);
            p.d_rows = p.d_rows_c = nullptr;
            p.rows_cap = 0;
            FC_CUDA(cudaMalloc(&p.d_rows, sizeof(int) * rows.size()));
            FC_CUDA(cudaMalloc(&p.d_rows_c, sizeof(int) * rows.size()));
            p.rows_cap = rows.size();
        }
        FC_CUDA(cudaMemcpy(p.d_rows_c, rows_c.data(), sizeof(int) * rows_c.size(), cudaMemcpyHostToDevice));
        if (!p.d_planes) FC_CUDA(cudaMalloc(&p.d_planes, sizeof(int) * p.g.nz));
        if (!p.d_plane_mask) FC_CUDA(cudaMalloc(&p.d_plane_mask, (size_t)p.g.nz));
        FC_CUDA(cudaMemcpy(p.d_rows, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice));
        FC_CUDA(cudaMemcpy(p.d_planes, planes.data(), sizeof(int) * planes.size(), cudaMemcpyHostToDevice));
        FC_CUDA(cudaMemcpy(p.d_plane_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice));
        p.n_rows = (long long)rows.size();
        p.n_planes = (int)planes.size();
        p.h_planes = planes;
        p.psf_window_z0 = win_z0;
        p.psf_window_planes = win_planes;
        p.hwin_valid = false;
        // CSR tap lists over the processed rows: tap (a,b,c) -> flat position (reference placement,
        // src/convolution3Dfft.cu:145-164) -> (row, x)
        {
            std::vector<int> plane_slot((size_t)p.g.nz, -1);
            for (size_t i = 0; i < planes.size(); ++i) plane_slot[(size_t)planes[i]] = (int)i;
            const long long d0 = pdims[3], d1 = pdims[4], d2 = pdims[5];
            const int k0 = pdims[0], k1 = pdims[1], k2 = pdims[2];
            const size_t K = (size_t)k0 * k1 * k2;
            std::vector<int> start(rows.size() + 1, 0), tx(K), tidx(K), lrow(K);
            size_t t = 0;
            for (int a = 0; a < k0; ++a) {
                long long aq = a - k0 / 2;
                if (aq < 0) aq += d0;
                for (int b = 0; b < k1; ++b) {
                    long long bq = b - k1 / 2;
                    if (bq < 0) bq += d1;
                    for (int c = 0; c < k2; ++c, ++t) {
                        long long cq = c - k2 / 2;
                        if (cq < 0) cq += d2;
                        const long long flat = cq + d2 * (bq + d1 * aq);
                        const long long grow = flat / p.g.nx;
                        const int slot = plane_slot[(size_t)(grow / p.g.ny)];
                        lrow[t] = row_index[(size_t)slot * p.g.ny + (size_t)(grow % p.g.ny)];
                        start[(size_t)lrow[t] + 1]++;
                    }
                }
            }
            for (size_t i = 0; i < rows.size(); ++i) start[i + 1] += start[i];
            std::vector<int> fill(start.begin(), start.end() - 1);
            t = 0;
            for (int a = 0; a < k0; ++a) {
                long long aq = a - k0 / 2;
                if (aq < 0) aq += d0;
                for (int b = 0; b < k1; ++b) {
                    long long bq = b - k1 / 2;
                    if (bq < 0) bq += d1;
                    for (int c = 0; c < k2; ++c, ++t) {
                        long long cq = c - k2 / 2;
                        if (cq < 0) cq += d2;
                        const long long flat = cq + d2 * (bq + d1 * aq);
                        const int at = fill[(size_t)lrow[t]]++;
                        tx[(size_t)at] = (int)(flat % p.g.nx);
                        tidx[(size_t)at] = (int)t;
                    }
                }
            }
            cudaFree(p.d_tap_start);
            p.d_tap_start = nullptr;
            FC_CUDA(cudaMalloc(&p.d_tap_start, sizeof(int) * start.size()));
            if (K > p.taps_cap) {
                cudaFree(p.d_tap_x);
                cudaFree(p.d_tap_idx);
                p.d_tap_x = p.d_tap_idx = nullptr;
                p.taps_cap = 0;
                FC_CUDA(cudaMalloc(&p.d_tap_x, sizeof(int) * K));
                FC_CUDA(cudaMalloc(&p.d_tap_idx, sizeof(int) * K));
                p.taps_cap = K;
            }
            FC_CUDA(cudaMemcpy(p.d_tap_start, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice));
            FC_CUDA(cudaMemcpy(p.d_tap_x, tx.data(), sizeof(int) * K, cudaMemcpyHostToDevice));
            FC_CUDA(cudaMemcpy(p.d_tap_idx, tidx.data(), sizeof(int) * K, cudaMemcpyHostToDevice));
        }
        std::memcpy(p.psf_key, pdims, sizeof(int) * 6);
    }
}

// zero the planes of `spec` ([nz][ny][xcp]) that the PSF passes work on, one memset per run of consecutive planes
static void clear_listed_planes(ConvPlan& p, float2* spec, cudaStream_t st)
{
    const size_t splane = (size_t)p.g.ny * p.g.xcp;
    for (int i = 0; i < p.n_planes;) {
        int j = i + 1;
        while (j < p.n_planes && p.h_planes[(size_t)j] == p.h_planes[(size_t)j - 1] + 1) ++j;
        FC_CUDA(cudaMemsetAsync(spec + (size_t)p.h_planes[(size_t)i] * splane, 0, (size_t)(j - i) * splane * sizeof(float2), st));
        i = j;
    }
}

void ensure_full_workspace(ConvPlan& p)
{
    if (!p.d_H) FC_CUDA(device_alloc_retry(&p.d_H, p.spec_bytes()));
}

void run_psf_spectrum(ConvPlan& p, const float* d_kernel, const int* pdims, cudaStream_t st)
{
    p.h_valid = false;   // the caller re-validates the cache when it knows the taps (host-pointer calls)
    ensure_full_workspace(p);
    psf_lists(p, pdims, st);
    clear_listed_planes(p, p.d_H, st);   // rows without taps are not transformed (psf_lists)
    XArgs xa = x_args(p);
    xa.spec = p.d_H;
    xa.nrows = p.n_rows;
    xa.rowList = p.d_rows;
    xa.psf.kernel = d_kernel;
    xa.psf.k0 = pdims[0];
    xa.psf.k1 = pdims[1];
    xa.psf.k2 = pdims[2];
    xa.psf.d0 = pdims[3];
    xa.psf.d1 = pdims[4];
    xa.psf.d2 = pdims[5];
    xa.tapStart = p.d_tap_start;
    xa.tapX = p.d_tap_x;
    xa.tapIdx = p.d_tap_idx;
    {
        PassTimer t(kPassPsfX, st);
        launch_x_fwd(xa, true, st);
    }
    {
        PassTimer t(kPassPsfY, st);
        ColArgs ya = y_args(p, p.d_H);
        ya.groupList = p.d_planes;
        col_pass(ya, 0, p.n_planes, st);
    }
    {
        PassTimer t(kPassPsfZ, st);
        ColArgs za = z_args(p, p.d_H);
        za.rowMask = p.d_plane_mask;
        if (!(p.psf_window_z0 >= 0 && launch_psf_z_pruned(za, p.psf_window_z0, p.psf_window_planes, st)))
            col_pass(za, 0, 1, st);
    }
    count_launches(3);
}

static bool otf_launch(ConvPlan& p, ColArgs& za, cudaStream_t st, bool probe)
{
    za.winPlanes = p.psf_window_planes;
    if (launch_col_otf_tma(za, 1, p.psf_window_z0, st, probe)) return true;
    return p.psf_window_planes == 16 && launch_col_otf(za, 1, p.psf_window_z0, st, probe);
}

bool psf_window_applies(ConvPlan& p, const int* pdims, cudaStream_t st)
{
    psf_lists(p, pdims, st);
    if (p.psf_window_z0 < 0) return false;
    ColArgs probe = z_args(p, p.d_spec);
    return otf_launch(p, probe, st, true);
}

bool run_psf_window(ConvPlan& p, const float* d_kernel, const int* pdims, cudaStream_t st)
{
    if (!psf_window_applies(p, pdims, st)) return false;
    const size_t need = (size_t)p.psf_window_planes * p.g.ny * p.g.xcp;
    if (need > p.hwin_cap) {
        FC_CUDA(cudaStreamSynchronize(st));
        cudaFree(p.d_Hwin);
        p.d_Hwin = nullptr;
        p.hwin_cap = 0;
        FC_CUDA(device_alloc_retry(&p.d_Hwin, need * sizeof(float2)));
        p.hwin_cap = need;
    }
    p.hwin_valid = false;
    if (!p.d_win_slot) FC_CUDA(cudaMalloc(&p.d_win_slot, 16 * sizeof(int)));
    if (std::memcmp(p.win_key, pdims, sizeof(int) * 6) != 0) {
        int slots[16];
        for (int n = 0; n < 16; ++n) {
            const int z = (p.psf_window_z0 + n) % p.g.nz;
            slots[n] = -1;
            for (int i = 0; i < p.n_planes; ++i)
                if (p.h_planes[(size_t)i] == z) slots[n] = i;
        }
        FC_CUDA(cudaStreamSynchronize(st));
        FC_CUDA(cudaMemcpy(p.d_win_slot, slots, sizeof(slots), cudaMemcpyHostToDevice));
        std::memcpy(p.win_key, pdims, sizeof(int) * 6);
    }
    FC_CUDA(cudaMemsetAsync(p.d_Hwin, 0, need * sizeof(float2), st));   // rows without taps are not transformed
    XArgs xa = x_args(p);
    xa.spec = p.d_Hwin;
    xa.nrows = p.n_rows;
    xa.rowList = p.d_rows_c;   // output row in the compact buffer
    xa.psf.kernel = d_kernel;
    xa.psf.k0 = pdims[0];
    xa.psf.k1 = pdims[1];
    xa.psf.k2 = pdims[2];
    xa.psf.d0 = pdims[3];
    xa.psf.d1 = pdims[4];
    xa.psf.d2 = pdims[5];
    xa.tapStart = p.d_tap_start;
    xa.tapX = p.d_tap_x;
    xa.tapIdx = p.d_tap_idx;
    {
        PassTimer t(kPassPsfX, st);
        launch_x_fwd(xa, true, st);
    }
    {
        PassTimer t(kPassPsfY, st);
        col_pass(y_args(p, p.d_Hwin), 0, p.n_planes, st);
    }
    count_launches(2);
    return true;
}

// ---- the image path in three pieces; the x/y pieces work on any range of z planes so that the host-pointer
// ---- entry can overlap them with the upload / download of the neighbouring planes
void run_xy_forward_planes(ConvPlan& p, const float* d_real, int z0, int n, cudaStream_t st, const PadGeom* pad)
{
    const size_t rplane = (size_t)p.g.ny * p.g.nx, splane = (size_t)p.g.ny * p.g.xcp;
    XArgs xa = x_args(p);
    xa.in_real = d_real + z0 * rplane;
    xa.spec = p.d_spec + z0 * splane;
    xa.nrows = (long long)p.g.ny * n;
    if (pad != nullptr) {
        xa.in_real = d_real;   // the unpadded volume; the loader picks the source row of every padded row
        xa.padOn = 1;
        xa.padZ0 = z0;
        xa.pad = *pad;
    }
    {
        PassTimer t(kPassXFwd, st);
        launch_x_fwd(xa, false, st);
    }
    {
        PassTimer t(kPassYFwd, st);
        col_pass(y_args(p, p.d_spec + z0 * splane), 0, n, st);
    }
    count_launches(2);
}

void run_z_fused(ConvPlan& p, bool window, cudaStream_t st)
{
    ColArgs za = z_args(p, p.d_spec);
    // reference: scale = 1.0f/(float)(size_img), src/convolution3Dfft.cu:531
    za.scale = 1.0f / (float)((size_t)p.g.nx * (size_t)p.g.ny * (size_t)p.g.nz);
    PassTimer t(kPassZFused, st);
    if (window) {
        za.H = p.d_Hwin;
        za.winSlot = p.d_win_slot;
        if (!otf_launch(p, za, st, false))
            throw std::runtime_error("fcb200: internal error, on-the-fly z pass unavailable");
    } else {
        za.H = p.d_H;
        col_pass(za, 2, 1, st);
    }
    count_launches(1);
}

void run_yx_inverse_planes(ConvPlan& p, float* d_real, int z0, int n, cudaStream_t st, const PadGeom* pad)
{
    const size_t rplane = (size_t)p.g.ny * p.g.nx, splane = (size_t)p.g.ny * p.g.xcp;
    {
        PassTimer t(kPassYInv, st);
        col_pass(y_args(p, p.d_spec + z0 * splane), 1, n, st);
    }
    XArgs xa = x_args(p);
    xa.spec = p.d_spec + z0 * splane;
    xa.out_real = d_real + z0 * rplane;
    xa.nrows = (long long)p.g.ny * n;
    if (pad != nullptr) {
        xa.out_real = d_real;   // the unpadded volume; only interior rows / columns are stored
        xa.padOn = 1;
        xa.padZ0 = z0;
        xa.pad = *pad;
    }
    {
        PassTimer t(kPassXInv, st);
        launch_x_inv(xa, st);
    }
    count_launches(2);
}

void run_convolve_window(ConvPlan& p, float* d_real, cudaStream_t st)
{
    run_xy_forward_planes(p, d_real, 0, p.g.nz, st);
    run_z_fused(p, true, st);
    run_yx_inverse_planes(p, d_real, 0, p.g.nz, st);
}

void run_inverse(ConvPlan& p, float2* spec, float* d_real, cudaStream_t st)
{
    col_pass(z_args(p, spec), 1, 1, st);
    {
        PassTimer t(kPassYInv, st);
        col_pass(y_args(p, spec), 1, p.g.nz, st);
    }
    XArgs xa = x_args(p);
    xa.spec = spec;
    xa.out_real = d_real;
    {
        PassTimer t(kPassXInv, st);
        launch_x_inv(xa, st);
    }
    count_launches(3);
}

void run_convolve(ConvPlan& p, float* d_real, cudaStream_t st)
{
    run_xy_forward_planes(p, d_real, 0, p.g.nz, st);
    run_z_fused(p, false, st);
    run_yx_inverse_planes(p, d_real, 0, p.g.nz, st);
}

// ------------------------------------------------------------------------------------------------
// slab-decomposed single volume (multi-GPU)
// ------------------------------------------------------------------------------------------------
static void check_slab(const ConvPlan& p, int nzl, int nyl)
{
    // ragged slabs: the last rank may own fewer than nyl rows / nzp planes; the block pitch stays nyl / nzp
    if (nzl <= 0 || nzl > p.g.nz || nyl <= 0 || nyl > p.g.ny)
        throw std::runtime_error("fcb200: slab extents must be positive and not larger than the volume");
}

void run_slab_xy_forward(ConvPlan& p, const float* d_real, float2* zslab, float2* send, int nzl, int nyl,
                         cudaStream_t st, float2* const* peers, int rank, int nzp, int z0, int nz_run)
{
    check_slab(p, nzl, nyl);
    if (nzp < nzl) throw std::runtime_error("fcb200: slab block pitch nzp must be >= the local plane count");
    // planes [z0, z0 + nz_run) of the local slab only (host-pointer calls overlap the upload of the next planes)
    if (nz_run < 0) nz_run = nzl - z0;
    if (z0 < 0 || nz_run < 0 || z0 + nz_run > nzl) throw std::runtime_error("fcb200: slab plane range out of bounds");
    if (nz_run == 0) return;
    const size_t rplane = (size_t)p.g.ny * p.g.nx, splane = (size_t)p.g.ny * p.g.xcp;
    d_real += (size_t)z0 * rplane;
    zslab += (size_t)z0 * splane;
    if (send) send += (size_t)z0 * nyl * p.g.xcp;
    nzl = nz_run;
    XArgs xa = x_args(p);
    xa.in_real = d_real;
    xa.spec = zslab;
    xa.nrows = (long long)p.g.ny * nzl;
    {
        PassTimer t(kPassXFwd, st);
        launch_x_fwd(xa, false, st);
    }
    ColArgs ya = y_args(p, zslab);
    ya.split = send;
    ya.splitRows = nyl;
    ya.splitBlock = (long long)nzp * nyl * p.g.xcp;
    ya.splitGroup = (long long)nyl * p.g.xcp;
    if (peers) {   // store straight into the peers' y-slab buffers [P*nzp][nyl][xcp]: my planes start at rank*nzp
        ya.split = nullptr;
        ya.splitPeers = peers;
        ya.splitPeerOffset = ((long long)rank * nzp + z0) * nyl * p.g.xcp;
    }
    {
        PassTimer t(kPassYFwd, st);
        col_pass(ya, 0, nzl, st);
    }
    count_launches(2);
}

void run_slab_z_fused(ConvPlan& p, float2* yslab, const float2* Hslab, int nyl, cudaStream_t st,
                      float2* const* peers, int rank, int nzl, float2* const* in_peers)
{
    check_slab(p, 1, nyl);
    ColArgs za = z_args(p, yslab);
    const long long C = (long long)nyl * p.g.xcp;
    za.stride = C;
    za.tilesPerGroup = (int)((C + 2 * za.txp - 1) / (2 * za.txp));
    za.rowLen = (int)C;
    za.H = Hslab;
    za.scale = 1.0f / (float)((size_t)p.g.nx * (size_t)p.g.ny * (size_t)p.g.nz);
    if (peers) {   // output plane z goes to rank z / nzl, into its receive buffer [P][nzl][nyl][xcp] at block `rank`
        za.splitPeers = peers;
        za.splitRows = nzl;
        za.splitPeerOffset = (long long)rank * nzl * C;
        za.splitGroup = 0;
        za.splitInPeers = in_peers;   // pull exchange: input planes straight from the ranks that hold them
    } else if (in_peers) {
        throw std::runtime_error("fcb200: the pull exchange needs the peer form of the fused z pass");
    }
    PassTimer t(kPassZFused, st);
    col_pass(za, 2, 1, st);
    count_launches(1);
}

void run_slab_yx_inverse(ConvPlan& p, const float2* recv, float2* zslab, float* d_real, int nzl, int nyl,
                         cudaStream_t st, int nzp, int z0, int nz_run)
{
    check_slab(p, nzl, nyl);
    if (nzp < nzl) throw std::runtime_error("fcb200: slab block pitch nzp must be >= the local plane count");
    if (nz_run < 0) nz_run = nzl - z0;
    if (z0 < 0 || nz_run < 0 || z0 + nz_run > nzl) throw std::runtime_error("fcb200: slab plane range out of bounds");
    if (nz_run == 0) return;
    d_real += (size_t)z0 * p.g.ny * p.g.nx;
    zslab += (size_t)z0 * p.g.ny * p.g.xcp;
    recv += (size_t)z0 * nyl * p.g.xcp;
    nzl = nz_run;
    ColArgs ya = y_args(p, zslab);
    ya.split = const_cast<float2*>(recv);
    ya.splitRows = nyl;
    ya.splitBlock = (long long)nzp * nyl * p.g.xcp;
    ya.splitGroup = (long long)nyl * p.g.xcp;
    {
        PassTimer t(kPassYInv, st);
        col_pass(ya, 1, nzl, st);
    }
    XArgs xa = x_args(p);
    xa.spec = zslab;
    xa.out_real = d_real;
    xa.nrows = (long long)p.g.ny * nzl;
    {
        PassTimer t(kPassXInv, st);
        launch_x_inv(xa, st);
    }
    count_launches(2);
}

size_t psf_slab_scratch_elems(ConvPlan& p, const int* pdims)
{
    std::vector<unsigned char> mask;
    std::vector<int> planes;
    int z0 = -1, wp = 0;
    psf_plane_list(pdims, p.g, mask, planes, z0, wp);
    return planes.size() * (size_t)p.g.ny * p.g.xcp;
}

void run_slab_psf(ConvPlan& p, const float* d_kernel, const int* pdims, int y0, int nyl, float2* Hslab,
                  float2* scratch, cudaStream_t st)
{
    check_slab(p, 1, nyl);
    if (y0 < 0 || y0 >= p.g.ny) throw std::runtime_error("fcb200: PSF slab out of range");
    const int rows_here = std::min(nyl, p.g.ny - y0);   // ragged last slab: fewer valid rows, same pitch
    psf_lists(p, pdims, st);
    // x pass (gather loader) on the rows that hold taps, written into the compact scratch (cleared first)
    FC_CUDA(cudaMemsetAsync(scratch, 0, (size_t)p.n_planes * p.g.ny * p.g.xcp * sizeof(float2), st));
    XArgs xa = x_args(p);
    xa.spec = scratch;
    xa.nrows = p.n_rows;
    xa.rowList = p.d_rows_c;
    xa.psf.kernel = d_kernel;
    xa.psf.k0 = pdims[0];
    xa.psf.k1 = pdims[1];
    xa.psf.k2 = pdims[2];
    xa.psf.d0 = pdims[3];
    xa.psf.d1 = pdims[4];
    xa.psf.d2 = pdims[5];
    xa.tapStart = p.d_tap_start;
    xa.tapX = p.d_tap_x;
    xa.tapIdx = p.d_tap_idx;
    {
        PassTimer t(kPassPsfX, st);
        launch_x_fwd(xa, true, st);
    }
    {
        PassTimer t(kPassPsfY, st);
        col_pass(y_args(p, scratch), 0, p.n_planes, st);
    }
    // rows [y0, y0+nyl) of every active plane -> its place in the y-slab; the z pass skips other planes
    // (one 2D copy per run of consecutive planes: a PSF occupies at most two runs, [0, k/2] and [nz - k/2, nz))
    const size_t row_bytes = (size_t)rows_here * p.g.xcp * sizeof(float2);
    for (int i = 0; i < p.n_planes;) {
        int j = i + 1;
        while (j < p.n_planes && p.h_planes[(size_t)j] == p.h_planes[(size_t)j - 1] + 1) ++j;
        const int z = p.h_planes[(size_t)i];
        FC_CUDA(cudaMemcpy2DAsync(Hslab + (size_t)z * nyl * p.g.xcp, (size_t)nyl * p.g.xcp * sizeof(float2),
                                  scratch + ((size_t)i * p.g.ny + y0) * p.g.xcp, (size_t)p.g.ny * p.g.xcp * sizeof(float2),
                                  row_bytes, (size_t)(j - i), cudaMemcpyDeviceToDevice, st));
        i = j;
    }
    ColArgs za = z_args(p, Hslab);
    const long long C = (long long)nyl * p.g.xcp;
    za.stride = C;
    za.tilesPerGroup = (int)((C + 2 * za.txp - 1) / (2 * za.txp));
    za.rowLen = (int)C;
    za.rowMask = p.d_plane_mask;
    {
        PassTimer t(kPassPsfZ, st);
        if (!(p.psf_window_z0 >= 0 && launch_psf_z_pruned(za, p.psf_window_z0, p.psf_window_planes, st)))
            col_pass(za, 0, 1, st);
    }
    count_launches(3);
}

}  // namespace fcb200
