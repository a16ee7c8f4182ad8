// gs_fused.cu — the whole path  cloud -> range image -> angle image -> ground mask -> cloud  as ONE persistent kernel.
//
// A thread-block cluster owns one frame at a time; CTA `rank` of the cluster owns the column slab
// [rank*CW, rank*CW + CW) of the range image and keeps it in shared memory from the projection to the
// back-projection, so that per frame HBM sees exactly the algorithmic bytes: 16 B per input point, 2 B per pixel of
// the no-ground image, 16 B per surviving point.  Nothing in between (pixel owners, range image, angle
// image, label image, dilation, row counts) ever leaves the SMs.
//
// Stages inside the kernel (reference file:line):
//   P  projection            Alfa_GS::CreateRangeImage  src/alfa_gs.cpp:247-298, lib/utils.h:8-43  (variant:
//                            Cloud2RangeNode::CloudCb src/cloud2range_node.cpp:396-455)
//   A  angle image           Alfa_GS::CreateAngleImg    src/alfa_gs.cpp:335-386
//   G  ground labelling      Alfa_GS::EraseBFS seeds :720-739 + Label::LabelOneComponent Labeler/label.cpp:48-104,
//                            restated as reachability in a union-find forest over vertical runs
//   D  dilate + erase        GetUniformKernel + cv::dilate + erase  src/alfa_gs.cpp:747-764
//   B  back-projection       Alfa_GS::CameraCb  src/alfa_gs.cpp:997-1074 (variant: range2cloud_node.cpp:64-91)
//
// Pixel ownership ("the last point in input order wins", alfa_gs.cpp:291): one 32-bit shared-memory max per point on the
// key (order tag << 16 | encoded range), order tag = (index >> s) + 1 with the smallest s that fits 16 bits; points that
// share a tag are adjacent lanes and are resolved with shuffles first.  Blocks of 256 points are dealt to the CTA predicted
// (from 4 sampled points) to own their columns, so for ordered clouds (sensor firing order or row-major organised) almost
// every pixel is CTA-local; the rest go through a small remote list that is applied over distributed shared memory after
// the local rounds.  DESIGN.md 4.1 describes every stage and what was measured on the way.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <math_constants.h>

#include "gs_internal.h"
#include "gs_math.cuh"

namespace cg = cooperative_groups;

namespace gs {

namespace {

#ifndef GS_EXP_DOUBLE
#define GS_EXP_DOUBLE 0  // timing experiments: run one (idempotent) phase twice — 1 projection, 2 angle walk, 3 output loop, 4 forest
                         // read-out, 5 links + runs + seeds, 7 link scan (the label stage is NOT idempotent: it rewrites its seed mask); the difference to the
                         // normal build, timed on the same box (tools/ab_bench.sh), is what the phase costs in situ
#endif
#define GS_REPEAT(k) for (int rep_ = 0; rep_ < (GS_EXP_DOUBLE == (k) ? 2 : 1); ++rep_)
#ifndef GS_READOUT_COMPRESS
#define GS_READOUT_COMPRESS 1
#endif
constexpr bool READOUT_COMPRESS = GS_READOUT_COMPRESS != 0;

constexpr int BLK = 256;        // points per dealt block
constexpr int MAX_BLOCKS = 4096;  // per frame: up to 1 Mi points
constexpr int MAX_CL = 8;
constexpr int MAX_H = 128;

// constants of the cold classification path (project_undecided), in device memory behind the KParams copy
struct GateConsts {
    double range_lo_in, range_hi_in, range_lo_out, range_hi_out;  // range gates -/+ 1e-9 m
    float alt_lo, alt_hi;                                         // altitude gates as floats
};

struct FusedArgs {
    KParams kp;
    Luts lut;
    int CL, CW;   // CTAs per frame, columns per CTA (the last CTA may own fewer)
    int S16;      // row stride of the range / tag slabs, u16 elements
    int S8;       // row stride of the angle slab, bytes
    int NG;       // 32-column groups per slab
    // shared memory offsets (bytes)
    unsigned off_range, off_tag, off_angle, off_masks, off_misc, off_lut;
    const float4* points;
    const long long* offsets;
    int frames;
    uint16_t* range_out;
    uint16_t* angle_out;
    uint16_t* res_out;
    float4* xyz_out;
    unsigned* counts_out;
    Counters* cnt;
    int stats;
    uint2* rlist;        // [clusters][CL][rcap] remote records
    unsigned* rcount;    // [clusters][CL]
    unsigned rcap;
    // ---- fast (float / fixed-point) projection, certified by margins (see project_lean2)
    double enc_scale;          // 65534 / (max_range - min_range)
    float alt_to_fix;          // altitude (rad) -> row cells * 2^FB
    float az_to_fix;           // angle (rad)    -> column cells * 2^FB
    float alt_lo, alt_hi;      // altitude gates as floats
    int row0_fix;              // rowfix = row0_fix - altitude_fix
    int quarter_fix;           // a quarter turn in column cells * 2^FB
    int col_half_fix;          // + half a cell (alfa_gs variant only)
    int row_margin, col_margin;  // fixed-point error margins
    unsigned cw_magic;         // ceil(2^32 / CW): owner = (col * cw_magic) >> 32 for col < 2^16
    unsigned ng_magic;         // ceil(2^32 / NG)
    const KParams* kp_dev;     // device-memory copy of kp for the out-of-line exact path
    const GateConsts* gates_dev;  // device-memory constants of project_undecided
    int col_limit, col_wrap;   // largest column the fast path may return; GS_COL_WRAP
    double range_lo_in, range_hi_in, range_lo_out, range_hi_out;  // range gates -/+ 1e-9 m
    // lean fast path (project_lean2): encoded range in 2^-UB fixed point, altitude gates in cell fixed point
    double u_scale, u_c0;      // ufix = trunc(range * u_scale + u_c0) = ((range - min_range) * 65534 / span + 1) * 2^UB
    int alt_gate_lo, alt_gate_span;  // rn(altitude * alt_to_fix) - alt_gate_lo <= alt_gate_span (unsigned): certainly inside the altitude gates
    float sg_t[11];            // Savitzky-Golay taps (cloud2range_node.cpp:470), sg_ws = 0: no smoothing
    int sg_ws;
    int repair;                // typed RepairGaps(range, 5, 1.0f) feeds the angle image (cloud2range_node.cpp:468-469)
    int use_fast;
    // frame queue shared by every launch that works on this batch: a cluster starts with frame first_frame + its index and
    // fetches every further frame with an atomic on *queue (zeroed by the host before the launches): frame = queue_base + old value
    unsigned* queue;
    int first_frame, queue_base;
    int slot_base;             // this launch's clusters use the scratch slots slot_base, slot_base + 1, ... (rlist / rcount / mail)
    struct SoftMail* mail;     // software clusters only: one mailbox per cluster slot
};

// Mailbox of one SOFTWARE cluster: CL ordinary CTAs that own the column slabs of one frame like the CTAs of a thread-block
// cluster do, but exchange their boundary data through global memory (L2) instead of distributed shared memory, and meet
// at a counter instead of barrier.cluster.  Hardware clusters must sit inside one GPC, which strands SMs (clusters of 4 reach
// 132 of a B200's 148 SMs); software clusters take whatever SMs are left.  Every CTA writes only its own slots.
struct SoftMail {
    unsigned bar;                                   // arrivals since the launch started
    unsigned next_frame;                            // published by rank 0 once per frame
    unsigned pad[30];
    unsigned chg[2][MAX_CL];                        // boundary exchange: "something changed in my slab" per parity and rank
    unsigned pub[MAX_CL][2][2][4];                  // [rank][parity][column 0 / last column][word]: ground masks of the boundary columns
    unsigned valid0[MAX_CL][4], validL[MAX_CL][4];  // VALID words of every CTA's column 0 / last column
    unsigned rowcnt[MAX_CL][MAX_H];                 // surviving pixels per row and slab
    unsigned char angle0[MAX_CL][MAX_H], angleL[MAX_CL][MAX_H];  // angles of every CTA's column 0 / last column
};

constexpr int FB = 18;  // fractional bits of the fixed-point cell coordinates


// Shared-memory layout of one CTA: a pure function of (H, CW, NW) so that it can be evaluated at compile time for the
// specialised instantiations of k_frame (constant strides and offsets fold into the addressing) and at run time otherwise.
struct Layout {
    int CL, CW, S16, S8, NG;
    unsigned off_range, off_tag, off_angle, off_masks, off_misc, off_lut;
    size_t total;
    bool ok;
};
constexpr int MK_KINDS = 8;  // == MK_COUNT (static_assert below)
__host__ __device__ constexpr size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ constexpr size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }
__host__ __device__ constexpr Layout layout_hcw(int H, int CW, int NW) {
    Layout L{};
    L.CL = 0;
    L.CW = CW;
    L.S16 = CW + 2 + ((CW & 1) ? 1 : 0);  // odd word stride: a column-major walk over rows hits distinct banks
    if (((L.S16 / 2) & 1) == 0) L.S16 += 2;
    L.S8 = static_cast<int>(align_up(static_cast<size_t>(CW) + 4, 4));
    if (((L.S8 / 4) & 1) == 0) L.S8 += 4;
    L.NG = (CW + 31) / 32;
    L.ok = static_cast<size_t>(H) * L.S16 <= 65535   // slab addresses are 16-bit
           && static_cast<size_t>(H) * CW <= 32768;  // run indices are 15-bit
    const size_t range_bytes = align_up(static_cast<size_t>(H) * L.S16 * 2, 16);
    const size_t tag_bytes = align_up(max_sz(range_bytes, static_cast<size_t>(H) * CW * 2), 16);
    const size_t angle_bytes = align_up(max_sz(static_cast<size_t>(H) * L.S8, static_cast<size_t>(4096) * 3), 16);  // 4096 = MAX_BLOCKS
    const size_t masks_bytes = align_up(static_cast<size_t>(MK_KINDS) * NW * CW * 8 + static_cast<size_t>(2) * H * L.NG * 4, 16);
    const size_t Hp = (static_cast<size_t>(H) + 7) & ~static_cast<size_t>(7);
    const size_t misc_bytes = align_up(14 * 8 + static_cast<size_t>(8) * Hp * 4 + Hp * 4 + Hp + 2 * 8 * 4 + (8 + 20 + 64) * 4 + 64 + Hp + 16, 16);  // 8 = MAX_CL
    L.off_range = 0;
    L.off_tag = static_cast<unsigned>(range_bytes);
    L.off_angle = static_cast<unsigned>(range_bytes + tag_bytes);
    L.off_masks = static_cast<unsigned>(range_bytes + tag_bytes + angle_bytes);
    L.off_misc = static_cast<unsigned>(range_bytes + tag_bytes + angle_bytes + masks_bytes);
    // per-CTA copies of the lookup tables: cos/sin(azimuth) of my columns, cos/sin(altitude), sin/cos(angle_y) as double and float
    const size_t lut_bytes = align_up(static_cast<size_t>(2) * CW * 8 + static_cast<size_t>(4) * H * 8 + static_cast<size_t>(2) * H * 4, 16);
    L.off_lut = static_cast<unsigned>(range_bytes + tag_bytes + angle_bytes + masks_bytes + misc_bytes);
    L.total = range_bytes + tag_bytes + angle_bytes + masks_bytes + misc_bytes + lut_bytes;
    return L;
}



// ---- small helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ double decode_range_d(const KParams& kp, unsigned enc) {
    // static_cast<double>(enc - 1) / 65534 * (max_range - min_range) + min_range   (alfa_gs.cpp:359-360, :1042-1045)
    // The division by the constant 65534 is done as q0 = n * R, r = fma(-q0, 65534, n), q = fma(r, R, q0) with
    // R = RN(1 / 65534): correctly rounded for every n in [0, 65534] (checked exhaustively against n / 65534.0 by
    // tests/clone_check.cpp), 3 FP64 instructions instead of the ~25 of a generic double division.
    // (double)(enc - 1) without the conversion unit (I2F.F64 runs on the quarter-rate XU pipe, which the angle walk and the
    // output saturate): 2^52 + n is exact, so is the subtraction
    const double n = __dsub_rn(__hiloint2double(0x43300000, static_cast<int>(enc) - 1), 4503599627370496.0);
    const double R = 1.0 / 65534.0;
    const double q0 = __dmul_rn(n, R);
    const double r = __fma_rn(-q0, 65534.0, n);
    const double norm = __fma_rn(r, R, q0);
    return __dadd_rn(__dmul_rn(norm, kp.span), kp.min_range);
}

// the bit-exact atan2f clone, out of line: it runs for ~1e-3 of the pixels
__device__ __noinline__ float atan2f_exact(float y, float x) { return gs_atan2f(y, x); }

// atan(a) for a in [0, 1]: a + a*s*q(s), s = a*a; max error 6.6e-8 rad in binary32 FMA arithmetic (fitted and measured
// by tools/fit_atan.py).  Used only on the certified fast paths: every result that lands within a margin >= 4x the error
// budget of a quantisation edge is recomputed with the reference's exact arithmetic.
__device__ __forceinline__ float atan01(float a) {
    const float s = __fmul_rn(a, a);
    float q = 0.0021047850605100393f;
    q = __fmaf_rn(q, s, -0.012938526459038258f);
    q = __fmaf_rn(q, s, 0.03731348365545273f);
    q = __fmaf_rn(q, s, -0.07019482553005219f);
    q = __fmaf_rn(q, s, 0.1039692610502243f);
    q = __fmaf_rn(q, s, -0.14136579632759094f);
    q = __fmaf_rn(q, s, 0.19983631372451782f);
    q = __fmaf_rn(q, s, -0.3333265483379364f);
    return __fmaf_rn(__fmul_rn(a, s), q, a);
}

// ---- sm_100 packed binary32 arithmetic (FFMA2 / FMUL2 / FADD2): one instruction, two independent IEEE operations — each half
// rounds exactly like the scalar fma.rn / mul.rn / add.rn, so a packed evaluation is bit-identical to two scalar ones and the
// certified margins carry over unchanged.  The kernel is bound by instruction issue, not by the FP32 pipe: pairing the two
// points of a projection unit (and two rows of an angle walk) halves the issue slots of their float chains.
__device__ __forceinline__ unsigned long long f2_bits(float2 v) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
    return r;
}
__device__ __forceinline__ float2 f2_from(unsigned long long r) {
    float2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
    return v;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return f2_from(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(r);
}
__device__ __forceinline__ float2 splat2(float c) { return make_float2(c, c); }
// atan01 of two arguments at once (the same operations in the same order as atan01: bit-identical per half)
__device__ __forceinline__ float2 atan01x2(float2 a) {
    const float2 s = mul2(a, a);
    float2 q = splat2(0.0021047850605100393f);
    q = fma2(q, s, splat2(-0.012938526459038258f));
    q = fma2(q, s, splat2(0.03731348365545273f));
    q = fma2(q, s, splat2(-0.07019482553005219f));
    q = fma2(q, s, splat2(0.1039692610502243f));
    q = fma2(q, s, splat2(-0.14136579632759094f));
    q = fma2(q, s, splat2(0.19983631372451782f));
    q = fma2(q, s, splat2(-0.3333265483379364f));
    return fma2(mul2(a, s), q, a);
}

// Union-find over vertical runs, 16-bit parents in shared memory: bits 0..14 = parent index, bit 15 = "ground" flag
// (meaningful on roots).  A root is always hung under a root of higher priority, priority = a multiplicative hash of the
// index (ties impossible: the hash is a bijection on 16 bits), so the forest is acyclic and — unlike linking by index,
// where the simultaneous unions of neighbouring columns build a chain as long as the slab is wide — its expected height
// is logarithmic (randomised linking); finds halve the paths they walk.
__device__ __forceinline__ unsigned uf16_prio(int x) { return (static_cast<unsigned>(x) * 40503u) & 0xffffu; }  // odd multiplier: bijective mod 2^16
__device__ __forceinline__ int uf16_find(volatile uint16_t* par, int x) {
    // path halving: only non-root entries are rewritten (plain 16-bit stores), roots change through atomicCAS / atomicOr only
    for (;;) {
        const int p = par[x] & 0x7fff;
        if (p == x) return x;
        const int gp = par[p] & 0x7fff;
        if (gp == p) return p;
        par[x] = static_cast<uint16_t>(gp);
        x = gp;
    }
}
// ---- explicit shared::cta accesses on 32-bit window addresses.  Plain C++ dereferences of pointers into the dynamic
// shared array compile to LDS/STS on shared::cluster addresses (this kernel maps neighbour CTAs' shared memory), and in
// divergent code the compiler re-derives the window base from SR_CgaCtaId (S2R) at every access.  The latency-bound
// union-find loops use these instead, on a base address that is kept in a register.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned lds_u16(unsigned a) {
    unsigned short v;
    asm volatile("ld.shared::cta.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
    unsigned v;
    asm volatile("ld.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32(unsigned a) {
    float v;
    asm volatile("ld.shared::cta.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned a) {
    double v;
    asm volatile("ld.shared::cta.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(unsigned a, unsigned v) {
    asm volatile("st.shared::cta.u16 [%0], %1;" ::"r"(a), "h"(static_cast<unsigned short>(v)) : "memory");
}
__device__ __forceinline__ unsigned cas_shared_u32(unsigned a, unsigned expect, unsigned want) {
    unsigned old;
    asm volatile("atom.shared::cta.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(expect), "r"(want) : "memory");
    return old;
}

// uf16_* on a forest addressed by its shared-window byte address `par`
__device__ __forceinline__ int uf16s_find(unsigned par, int x) {
    for (;;) {
        const int p = lds_u16(par + 2 * x) & 0x7fff;
        if (p == x) return x;
        const int gp = lds_u16(par + 2 * p) & 0x7fff;
        if (gp == p) return p;
        sts_u16(par + 2 * x, gp);
        x = gp;
    }
}
// roots of two nodes found in lock-step (independent walks: one shared-memory round trip per hop for both), no writes
__device__ __forceinline__ void uf16s_roots2(unsigned par, int& a, int& b) {
    int pa = lds_u16(par + 2 * a) & 0x7fff, pb = lds_u16(par + 2 * b) & 0x7fff;
    while (pa != a || pb != b) {
        a = pa;
        b = pb;
        pa = lds_u16(par + 2 * a) & 0x7fff;
        pb = lds_u16(par + 2 * b) & 0x7fff;
    }
}
// same walk; returns the raw root entries (bit 15 = ground flag)
__device__ __forceinline__ void uf16s_root_entries2(unsigned par, int a, int b, unsigned& ea, unsigned& eb) {
    ea = lds_u16(par + 2 * a);
    eb = lds_u16(par + 2 * b);
    while (static_cast<int>(ea & 0x7fff) != a || static_cast<int>(eb & 0x7fff) != b) {
        a = ea & 0x7fff;
        b = eb & 0x7fff;
        ea = lds_u16(par + 2 * a);
        eb = lds_u16(par + 2 * b);
    }
}
// Concurrent union with a compare-and-swap (any thread of the CTA may work on any part of the forest).  The node whose
// parent entry is rewritten is verified to be a root by the compare-and-swap itself and is always hung under a node of
// higher priority, so the forest stays acyclic even when the other node has stopped being a root meanwhile.
__device__ __forceinline__ void uf16s_unite_cas(unsigned par, int a, int b) {
    uf16s_roots2(par, a, b);
    for (;;) {
        if (a == b) return;
        if (uf16_prio(a) > uf16_prio(b)) { const int t = a; a = b; b = t; }  // a: lower priority, hangs under b
        const unsigned w = par + 4 * (a >> 1);
        const int sh = (a & 1) * 16;
        const unsigned old = lds_u32(w);
        if (static_cast<int>((old >> sh) & 0xffffu) == a) {
            const unsigned want = (old & ~(0xffffu << sh)) | (static_cast<unsigned>(b) << sh);
            if (cas_shared_u32(w, old, want) == old) return;
            continue;  // the word changed (maybe only its other half): look again
        }
        a = uf16s_find(par, a);  // a stopped being a root
        b = uf16s_find(par, b);
    }
}
// Warp-synchronous union of one link per lane WITHOUT atomics, for links whose components no other warp touches (shared-
// memory atomics cost ~4 cycles per LANE, SM-wide: a compare-and-swap per link was two thirds of the union stage).
// Round: an active lane finds its two roots (reads only) and, if they differ, hangs the root of lower priority under the
// other with a plain 16-bit store; __syncwarp; the lane then checks that its store survived (another lane may have hung
// the same root elsewhere, or overwritten it meanwhile) and retries if not.  Parents always have a higher priority than
// their children, so the forest stays acyclic whatever the interleaving; a hook can only be lost to another store of the
// same round, which its owner detects after the barrier — links found united rest on hooks whose owners make them stick.
// There is no path compression here: a shortcut taken through a hook that is lost later would bypass a subtree.
__device__ __forceinline__ void uf16s_unite_warp(unsigned par, bool act, int a, int b) {
    for (;;) {
        int lo = 0, hi = 0;
        if (act) {
            uf16s_roots2(par, a, b);
            act = a != b;
            if (act) {
                const bool a_high = uf16_prio(a) > uf16_prio(b);
                lo = a_high ? b : a;
                hi = a_high ? a : b;
                sts_u16(par + 2 * lo, hi);
            }
        }
        if (!__any_sync(0xffffffffu, act)) return;
        __syncwarp();
        if (act) {
            act = (lds_u16(par + 2 * lo) & 0x7fff) != static_cast<unsigned>(hi);
            a = lo;
            b = hi;
        }
        if (!__any_sync(0xffffffffu, act)) return;
    }
}
__device__ __forceinline__ bool uf16_flag_root(uint16_t* par, int x) {  // returns true when the flag was newly set
    for (;;) {
        const int root = uf16_find(par, x);
        unsigned* w = reinterpret_cast<unsigned*>(par) + (root >> 1);
        const int sh = (root & 1) * 16;
        const unsigned old = atomicOr(w, 0x8000u << sh);
        // the flag phase runs after all unions: root is still a root
        return ((old >> sh) & 0x8000u) == 0;
    }
}

// Row masks: bit r of the 128-bit mask <-> image row r.
struct M128 {
    unsigned long long lo, hi;
};
template <int NW>
struct Mask;
template <>
struct Mask<1> {
    unsigned long long w0;
    __device__ __forceinline__ static Mask zero() { return Mask{0ull}; }
    __device__ __forceinline__ bool any() const { return w0 != 0; }
    __device__ __forceinline__ Mask operator|(Mask o) const { return Mask{w0 | o.w0}; }
    __device__ __forceinline__ Mask operator&(Mask o) const { return Mask{w0 & o.w0}; }
    __device__ __forceinline__ Mask operator^(Mask o) const { return Mask{w0 ^ o.w0}; }
    __device__ __forceinline__ Mask operator~() const { return Mask{~w0}; }
    __device__ __forceinline__ bool operator!=(Mask o) const { return w0 != o.w0; }
    __device__ __forceinline__ Mask shl(int k) const { return Mask{k >= 64 ? 0ull : w0 << k}; }
    __device__ __forceinline__ Mask shr(int k) const { return Mask{k >= 64 ? 0ull : w0 >> k}; }
    __device__ __forceinline__ Mask add(Mask o) const { return Mask{w0 + o.w0}; }
    __device__ __forceinline__ Mask rev() const { return Mask{__brevll(w0)}; }
    __device__ __forceinline__ bool bit(int r) const { return (w0 >> r) & 1ull; }
    __device__ __forceinline__ void set(int r) { w0 |= 1ull << r; }
    __device__ __forceinline__ int top() const { return w0 ? 63 - __clzll(w0) : -1; }
    __device__ __forceinline__ int lowest() const { return __ffsll(static_cast<long long>(w0)) - 1; }
    __device__ __forceinline__ void clear_lowest() { w0 &= w0 - 1ull; }
    __device__ __forceinline__ unsigned word32(int i) const { return static_cast<unsigned>(w0 >> (32 * i)); }
    __device__ __forceinline__ static Mask low(int n) { return Mask{n >= 64 ? ~0ull : ((1ull << n) - 1ull)}; }
};
template <>
struct Mask<2> {
    unsigned long long w0, w1;
    __device__ __forceinline__ static Mask zero() { return Mask{0ull, 0ull}; }
    __device__ __forceinline__ bool any() const { return (w0 | w1) != 0; }
    __device__ __forceinline__ Mask operator|(Mask o) const { return Mask{w0 | o.w0, w1 | o.w1}; }
    __device__ __forceinline__ Mask operator&(Mask o) const { return Mask{w0 & o.w0, w1 & o.w1}; }
    __device__ __forceinline__ Mask operator^(Mask o) const { return Mask{w0 ^ o.w0, w1 ^ o.w1}; }
    __device__ __forceinline__ Mask operator~() const { return Mask{~w0, ~w1}; }
    __device__ __forceinline__ bool operator!=(Mask o) const { return w0 != o.w0 || w1 != o.w1; }
    __device__ __forceinline__ Mask shl(int k) const {
        if (k == 0) return *this;
        if (k >= 128) return zero();
        if (k >= 64) return Mask{0ull, w0 << (k - 64)};
        return Mask{w0 << k, (w1 << k) | (w0 >> (64 - k))};
    }
    __device__ __forceinline__ Mask shr(int k) const {
        if (k == 0) return *this;
        if (k >= 128) return zero();
        if (k >= 64) return Mask{w1 >> (k - 64), 0ull};
        return Mask{(w0 >> k) | (w1 << (64 - k)), w1 >> k};
    }
    __device__ __forceinline__ Mask add(Mask o) const {
        const unsigned long long lo = w0 + o.w0;
        return Mask{lo, w1 + o.w1 + (lo < w0 ? 1ull : 0ull)};
    }
    __device__ __forceinline__ Mask rev() const { return Mask{__brevll(w1), __brevll(w0)}; }
    __device__ __forceinline__ bool bit(int r) const { return ((r < 64 ? w0 >> r : w1 >> (r - 64)) & 1ull) != 0; }
    __device__ __forceinline__ void set(int r) {
        if (r < 64) w0 |= 1ull << r;
        else w1 |= 1ull << (r - 64);
    }
    __device__ __forceinline__ int top() const { return w1 ? 127 - __clzll(w1) : (w0 ? 63 - __clzll(w0) : -1); }
    __device__ __forceinline__ int lowest() const { return w0 ? __ffsll(static_cast<long long>(w0)) - 1 : (w1 ? 63 + __ffsll(static_cast<long long>(w1)) : -1); }
    __device__ __forceinline__ void clear_lowest() {
        if (w0) w0 &= w0 - 1ull;
        else w1 &= w1 - 1ull;
    }
    __device__ __forceinline__ unsigned word32(int i) const { return static_cast<unsigned>((i < 2 ? w0 : w1) >> (32 * (i & 1))); }
    __device__ __forceinline__ static Mask low(int n) {
        if (n >= 128) return Mask{~0ull, ~0ull};
        if (n >= 64) return Mask{~0ull, n == 64 ? 0ull : ((1ull << (n - 64)) - 1ull)};
        return Mask{(1ull << n) - 1ull, 0ull};
    }
};

// All pixels joined to a seed in g through chains of set link bits: link bit r joins rows r and r+1.  The carry of
// one addition walks up a whole run of links ((A + c) ^ A), bit reversal gives the downward direction.
template <int NW>
__device__ __forceinline__ Mask<NW> fill_up(Mask<NW> g, Mask<NW> linked_to_prev) {
    const Mask<NW> c = g.shl(1) & linked_to_prev;
    return g | c | ((linked_to_prev.add(c) ^ linked_to_prev) & linked_to_prev);
}
template <int NW>
__device__ __forceinline__ Mask<NW> vfill(Mask<NW> g, Mask<NW> link, int H) {
    // upward (towards higher row index): pixel r is linked to its predecessor r-1 by link bit r-1
    const Mask<NW> up = fill_up<NW>(g, link.shl(1));
    // downward: reverse the bit order; pixel r's predecessor in reversed order is r+1, joined by link bit r.
    // rev() maps bit r -> (64*NW - 1 - r).
    const Mask<NW> dn = fill_up<NW>(g.rev(), link.rev()).rev();
    return (up | dn) & Mask<NW>::low(H);
}

// mask storage in shared memory: [kind][word][column], u64 words
enum { MK_VALID = 0, MK_VA, MK_HA, MK_V, MK_HM, MK_G, MK_KEEP, MK_S, MK_COUNT };
static_assert(MK_COUNT == MK_KINDS, "layout_hcw sizes the mask region for MK_KINDS kinds");

// Row masks live in shared memory as 32-bit words: m32[(kind * NWD + w) * CW + c] holds rows 32w .. 32w+31 of column c.
template <int NW>
__device__ __forceinline__ Mask<NW> ld_mask32(const unsigned* m32, int kind, int NWD, int CW, int c) {
    const unsigned* p = m32 + static_cast<size_t>(kind) * NWD * CW + c;
    Mask<NW> m;
    m.w0 = p[0];
    if (NWD > 1) m.w0 |= static_cast<unsigned long long>(p[CW]) << 32;
    if constexpr (NW == 2) {
        m.w1 = NWD > 2 ? p[2 * CW] : 0u;
        if (NWD > 3) m.w1 |= static_cast<unsigned long long>(p[3 * CW]) << 32;
    }
    return m;
}
template <int NW>
__device__ __forceinline__ void st_mask32(unsigned* m32, int kind, int NWD, int CW, int c, Mask<NW> m) {
    unsigned* p = m32 + static_cast<size_t>(kind) * NWD * CW + c;
    p[0] = static_cast<unsigned>(m.w0);
    if (NWD > 1) p[CW] = static_cast<unsigned>(m.w0 >> 32);
    if constexpr (NW == 2) {
        if (NWD > 2) p[2 * CW] = static_cast<unsigned>(m.w1);
        if (NWD > 3) p[3 * CW] = static_cast<unsigned>(m.w1 >> 32);
    }
}

// ---- projection of one point (double precision, the reference's operation order) ---------------------------------
struct Proj {
    int row, col;   // col may equal W for the alfa_gs variant (flat alias), row/col = -1 when dropped
    unsigned enc;
    bool near_edge;
};

// Rare path, kept out of line.  It reads the parameters from their device-memory copy: taking the address of the
// kernel's by-value argument would force a stack copy of the whole struct.
__device__ __noinline__ Proj project_point_d(const KParams* __restrict__ kpp, float fx, float fy, float fz) {
    const KParams& kp = *kpp;
    Proj out;
    out.row = -1;
    out.col = -1;
    out.enc = 0;
    out.near_edge = false;
    if (!(isfinite(fx) && isfinite(fy) && isfinite(fz))) return out;  // contract Q3
    const double x = fx, y = fy, z = fz;
    const double range = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));  // utils.h:41-43
    if (range < kp.min_range || range > kp.max_range) return out;                                             // alfa_gs.cpp:261
    const double a = atan2(y, x);
    const double azimuth = (y >= 0 ? a : __dadd_rn(a, 6.283185307179586476925286766559));  // utils.h:24-28
    const double altitude = atan2(z, hypot(x, y));                                        // utils.h:8-11
    if (altitude < kp.min_angle || altitude > kp.max_angle) return out;                   // alfa_gs.cpp:269
    double rowf, colf;
    if (kp.variant == GS_VARIANT_CLOUD2RANGE) {
        rowf = __dmul_rn(__ddiv_rn(__dsub_rn(kp.max_angle, altitude), kp.ang_span), static_cast<double>(kp.H));  // cloud2range_node.cpp:436
        colf = __ddiv_rn(azimuth, kp.d_az);                                                                      // :438
    } else {
        rowf = __dsub_rn(static_cast<double>(kp.H - 1),
                         __ddiv_rn(__dsub_rn(__dsub_rn(altitude, kp.half_d_alt), kp.min_angle), kp.d_alt));  // alfa_gs.cpp:275
        colf = __ddiv_rn(__dadd_rn(azimuth, kp.half_d_az), kp.d_az);                                         // :277
    }
    if (!(colf >= 0.0)) return out;  // contract Q3: y = -0.0, x < 0
    const int row = __double2int_rz(rowf);
    int col = __double2int_rz(colf);
    {
        const double fr = rowf - floor(rowf), fc = colf - floor(colf);
        const double cell_alt = (kp.variant == GS_VARIANT_CLOUD2RANGE) ? kp.ang_span / kp.H : kp.d_alt;
        out.near_edge = fmin(fr, 1.0 - fr) * cell_alt < 1e-5 || fmin(fc, 1.0 - fc) * kp.d_az < 1e-5;
    }
    if (row < 0 || row >= kp.H) return out;  // Q4
    if (kp.variant == GS_VARIANT_CLOUD2RANGE) {
        if (col >= kp.W) return out;
    } else if (kp.col_policy == GS_COL_WRAP) {
        col %= kp.W;
    } else {
        if (col > kp.W || (col == kp.W && row == kp.H - 1)) return out;  // Q2: flat index past the end of the image
    }
    const double range_norm = __ddiv_rn(__dsub_rn(range, kp.min_range), kp.span);  // alfa_gs.cpp:288
    out.enc = static_cast<unsigned>(__double2int_rz(__dadd_rn(__dmul_rn(range_norm, 65534.0), 1.0)));  // :291
    out.row = row;
    out.col = col;
    return out;
}

constexpr int UB = 14;  // fractional bits of the fixed-point encoded range

__device__ __forceinline__ float rsqrt_ftz(float a) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
__device__ __forceinline__ float rcp_ftz(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

// Certified fast projection: BRANCH-FREE float / fixed-point arithmetic whose every decision keeps a margin >= 4x its error
// budget; a point that lands inside a margin is "not certified" and goes to the reference's exact double arithmetic
// (project_point_d).  With GS_OPT_COLLECT_STATS the kernel evaluates both for every point and counts disagreements (0 in
// all tests).  Being straight-line code, the evaluations of two points interleave in the instruction stream, which hides
// the latency of the FP64 chain.
//  * range: double sum of squares, MUFU rsqrt seed of the float sum of squares (relative error e0 < 2.5e-7) and ONE Heron
//    step r + (s2 - r*r) * (0.5 * seed): relative error 1.5 e0^2 < 1e-13, i.e. < 1e-8 of an encoded-range unit (the
//    reference's sqrt(x*x + y*y + z*z), utils.h:41-43, differs from the true value by rounding only).  The encoded value
//    (alfa_gs.cpp:288-291) is taken in 2^-14 fixed point from one DFMA + F2I; a fractional part of 0 or 16383 (within
//    6.1e-5 of an integer) or a value outside (1, 65535) — the range gates of alfa_gs.cpp:261 map to exactly 1 and 65535 —
//    is not certified.
//  * altitude / azimuth (utils.h:8-28): float atan on [0, 1] (atan01, error 6.6e-8 rad); the azimuth after octant reduction
//    with the raw MUFU rcp (flush-to-zero form; operands below 1e-15 are not certified) and reflections in exact 18-bit fixed
//    point, the altitude from |z| * rsqrt(x^2 + y^2) for |altitude| <= 45 degrees (steeper points are not certified); the
//    altitude gates (alfa_gs.cpp:269) and the cell edges (:275, :277) are tested on the fixed-point cell coordinates with
//    margins of 1e-6 rad / 7.5e-7 rad.
// ok = pixel (row, col, enc) certified.  Everything else — outside a gate, within a margin of a gate or an edge, the
// flat-index alias column W, y == 0, non-finite — is classified by project_undecided / project_point_d.
struct Lean {
    int row, col;
    unsigned enc;
    bool ok;
};
// project_lean2 evaluates the two points a warp lane handles per iteration together: their float chains (sums of squares,
// quotients, both polynomial atans, scalings) run as packed f32x2 instructions, bit-identical to two scalar evaluations.
__device__ __forceinline__ void project_lean2(const FusedArgs& A, const float4& pa, const float4& pb, Lean& oa, Lean& ob) {
    const KParams& kp = A.kp;
    const float2 fx = make_float2(pa.x, pb.x), fy = make_float2(pa.y, pb.y), fz = make_float2(pa.z, pb.z);
    const float2 h2 = fma2(fx, fx, mul2(fy, fy));
    const float2 s2f = fma2(fz, fz, h2);
    const float2 hr = make_float2(rsqrt_ftz(h2.x), rsqrt_ftz(h2.y));
    // altitude = atan2(z, hypot(x, y)) for |altitude| <= 45 degrees: tan = |z| / hypot = |z| * rsqrt(x^2 + y^2), no reciprocal and
    // no octant reflection (relative error of the quotient: rsqrt.approx 2^-22.4 + the rounding of h2 and of one product, below
    // the reciprocal form's).  Steeper points are left to the exact path — lidar beams beyond +-45 degrees sit outside every
    // preset's altitude gates.  The azimuth octant and both points' polynomial chains run packed.
    const float2 az = make_float2(fabsf(fz.x), fabsf(fz.y));
    const float2 alt_q = mul2(az, hr);
    const float2 ax = make_float2(fabsf(fx.x), fabsf(fx.y)), ay = make_float2(fabsf(fy.x), fabsf(fy.y));
    const float2 azi_q = make_float2(__fmul_rn(fminf(ax.x, ay.x), rcp_ftz(fmaxf(ax.x, ay.x))), __fmul_rn(fminf(ax.y, ay.y), rcp_ftz(fmaxf(ax.y, ay.y))));
    const float2 tf = mul2(atan01x2(alt_q), splat2(A.alt_to_fix));  // |altitude| in cell fixed point; the sign is applied to the product (exact either way)
    const float2 ta = mul2(atan01x2(azi_q), splat2(A.az_to_fix));
    const unsigned rm = static_cast<unsigned>(A.row_margin), cm = static_cast<unsigned>(A.col_margin);
    auto finish = [&](float x, float y, float z, float h2v, float s2fv, float altq, float tfv, float tav, Lean& o) {
        const double xd = x, yd = y, zd = z;
        const double s2 = __fma_rn(xd, xd, __fma_rn(yd, yd, __dmul_rn(zd, zd)));
        const double rd = static_cast<double>(rsqrt_ftz(s2fv));
        const double r0 = __dmul_rn(s2, rd);
        const double range = __fma_rn(__fma_rn(-r0, r0, s2), __dmul_rn(0.5, rd), r0);
        const int ufix = __double2int_rz(__fma_rn(range, A.u_scale, A.u_c0));  // NaN -> 0, overflow saturates: both fail the test below
        const unsigned ufrac = static_cast<unsigned>(ufix) & ((1u << UB) - 1u);
        const bool range_ok = static_cast<unsigned>(ufix - ((1 << UB) + 1)) <= static_cast<unsigned>((65534 << UB) - 2) && ufrac - 1u <= (1u << UB) - 3u;
        const bool xy_ok = y != 0.0f && h2v > 1e-30f && altq <= 1.0f;  // false for NaN
        const int altfix = __float2int_rn(__uint_as_float(__float_as_uint(tfv) | (__float_as_uint(z) & 0x80000000u)));
        const int rowfix = A.row0_fix - altfix;
        const unsigned rfrac = static_cast<unsigned>(rowfix) & ((1u << FB) - 1u);
        int c = __float2int_rn(tav);
        if (fabsf(y) > fabsf(x)) c = A.quarter_fix - c;
        if (x < 0.0f) c = 2 * A.quarter_fix - c;
        if (y < 0.0f) c = 4 * A.quarter_fix - c;
        const int colfix = c + A.col_half_fix;
        const unsigned cfrac = static_cast<unsigned>(colfix) & ((1u << FB) - 1u);
        o.row = rowfix >> FB;
        o.col = colfix >> FB;
        o.enc = static_cast<unsigned>(ufix) >> UB;
        const bool gate_ok = static_cast<unsigned>(altfix - A.alt_gate_lo) <= static_cast<unsigned>(A.alt_gate_span);
        const bool edges_ok = rfrac - rm <= (1u << FB) - 2u * rm && cfrac - cm <= (1u << FB) - 2u * cm;
        const bool box_ok = static_cast<unsigned>(o.row) < static_cast<unsigned>(kp.H) && static_cast<unsigned>(o.col) < static_cast<unsigned>(kp.W);
        o.ok = range_ok && xy_ok && gate_ok && edges_ok && box_ok;
    };
    finish(pa.x, pa.y, pa.z, h2.x, s2f.x, alt_q.x, tf.x, ta.x, oa);
    finish(pb.x, pb.y, pb.z, h2.y, s2f.y, alt_q.y, tf.y, ta.y, ob);
}
// a point project_lean2 did not certify: 1 = certainly dropped by a gate (alfa_gs.cpp:261, :269), 2 = exact evaluation.
// Cold path: the range and the altitude are recomputed here (the same arithmetic as in project_lean2) rather than kept
// alive across the hot loop.
__device__ __noinline__ int project_undecided(const GateConsts* __restrict__ Ap, float fx, float fy, float fz) {
    const GateConsts& A = *Ap;
    const double x = fx, y = fy, z = fz;
    const double s2 = __fma_rn(x, x, __fma_rn(y, y, __dmul_rn(z, z)));
    const float h2 = __fmaf_rn(fx, fx, __fmul_rn(fy, fy));
    const double rd = static_cast<double>(rsqrt_ftz(__fmaf_rn(fz, fz, h2)));
    const double r0 = __dmul_rn(s2, rd);
    const double range = __fma_rn(__fma_rn(-r0, r0, s2), __dmul_rn(0.5, rd), r0);
    if (range < A.range_lo_out || range > A.range_hi_out) return 1;  // outside the range gates by more than 1e-9 m
    if (!(range > A.range_lo_in && range < A.range_hi_in)) return 2;  // within 1e-9 m of a gate, or not finite
    if (fy == 0.0f || !(h2 > 1e-30f)) return 2;                       // the float altitude is not meaningful
    const float az = fabsf(fz);
    const float h = __fmul_rn(h2, rsqrt_ftz(h2));
    float t = atan01(__fmul_rn(fminf(az, h), rcp_ftz(fmaxf(az, h))));
    if (az > h) t = 1.57079632679489662f - t;
    const float alt = __uint_as_float(__float_as_uint(t) | (__float_as_uint(fz) & 0x80000000u));
    if (alt < A.alt_lo - 1e-6f || alt > A.alt_hi + 1e-6f) return 1;   // outside the altitude gates by more than 1e-6 rad
    return 2;
}

// cheap column estimate used only to deal point blocks to CTAs (any answer is correct, a good one is fast)
__device__ __forceinline__ int approx_col(const FusedArgs& A, float x, float y) {
    // the azimuth of project_lean2 (octant reduction, polynomial atan, reflections in fixed point) without its certificates
    const float ax = fabsf(x), ay = fabsf(y);
    int c = __float2int_rn(__fmul_rn(atan01(__fmul_rn(fminf(ax, ay), rcp_ftz(fmaxf(fmaxf(ax, ay), 1e-30f)))), A.az_to_fix));
    if (ay > ax) c = A.quarter_fix - c;
    if (x < 0.0f) c = 2 * A.quarter_fix - c;
    if (y < 0.0f) c = 4 * A.quarter_fix - c;
    return min(max((c + A.col_half_fix) >> FB, 0), A.kp.W - 1);
}

// true when lane + d exists and holds the same value (SHFL.DOWN's in-range predicate: no lane arithmetic)
__device__ __forceinline__ bool shfl_down_eq(unsigned v, int d) {
    unsigned r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b32 o;\n\t"
        "shfl.sync.down.b32 o|p, %1, %2, 31, 0xffffffff;\n\t"
        "setp.eq.and.b32 q, o, %1, p;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(v), "r"(d));
    return r != 0;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// barrier of a software cluster: every CTA adds one to the mailbox counter and waits until all CL have (epoch counts the
// arrivals expected so far).  The CTAs of a software cluster are co-resident by construction (cooperative launch, one CTA per SM).
__device__ __forceinline__ void soft_cluster_sync(unsigned* bar, unsigned& epoch, int CL) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += static_cast<unsigned>(CL);
        __threadfence();  // my global stores (mailbox, remote records) before my arrival
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
        while (static_cast<int>(ld_acquire_gpu(bar) - epoch) < 0) {
        }
    }
    __syncthreads();
}

}  // namespace

// ====================================================================================================================
// the kernel
// ====================================================================================================================
// NW: 64-bit words per column mask (H <= 64 * NW); NT: threads per CTA (1024 x 1 CTA/SM or 512 x 2 CTAs/SM);
// HT, CWT, CLT: image rows, columns per CTA and CTAs per frame fixed at compile time (0 = read from the arguments; the image
// is then exactly CLT * CWT wide): the BASELINE sensor shapes get instantiations whose strides, offsets and trip counts are constants.
// SOFT: software clusters (see SoftMail) instead of thread-block clusters; launched without a cluster dimension.
template <int NW, int NT, int HT, int CWT, int CLT, bool SOFT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 1) k_frame(const FusedArgs A) {
    constexpr bool SPEC = HT > 0;
    constexpr Layout LC = layout_hcw(SPEC ? HT : 64, SPEC ? CWT : 512, NW);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const KParams& kp = A.kp;
    const int CL = SPEC ? CLT : A.CL, CW = SPEC ? CWT : A.CW, H = SPEC ? HT : kp.H, W = SPEC ? CLT * CWT : kp.W;
    const int rank = SOFT ? static_cast<int>(blockIdx.x % CL) : static_cast<int>(cluster.block_rank());
    const int cluster_id = blockIdx.x / CL;
    const int slot = A.slot_base + cluster_id;  // scratch slot of this cluster (remote lists, mailbox)
    SoftMail* const mail = SOFT ? A.mail + slot : nullptr;
    unsigned soft_epoch = 0;
    auto csync = [&]() {
        if constexpr (SOFT) soft_cluster_sync(&mail->bar, soft_epoch, CL);
        else cluster.sync();
    };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = rank * CW;                 // first global column of my slab
    const int myCW = SPEC ? CWT : min(CW, W - col0);  // columns I own (> 0 by construction)
    const int S16 = SPEC ? LC.S16 : A.S16, S8 = SPEC ? LC.S8 : A.S8, NG = SPEC ? LC.NG : A.NG;
    const unsigned off_range = SPEC ? LC.off_range : A.off_range, off_tag = SPEC ? LC.off_tag : A.off_tag, off_angle = SPEC ? LC.off_angle : A.off_angle;
    const unsigned off_masks = SPEC ? LC.off_masks : A.off_masks, off_misc = SPEC ? LC.off_misc : A.off_misc, off_lut = SPEC ? LC.off_lut : A.off_lut;
    // forest index of the run that starts at row rs of local column c.  The row is rotated by 2c inside the column so that
    // runs starting at the same row in neighbouring columns (the ground does) fall into different shared-memory banks.
    auto run_node = [=](int c, int rs) -> int { return c * H + ((rs + 2 * c) & (H - 1)); };  // H is a power of two (checked on the host)

    // shared-window address of the dynamic shared array, laundered through an opaque move so that it stays in a register
    unsigned smem_base32;
    {
        const unsigned t = smem_addr(smem_raw);
        asm volatile("mov.u32 %0, %1;" : "=r"(smem_base32) : "r"(t));
    }
    // ---- shared memory carve-up
    uint16_t* s_range = reinterpret_cast<uint16_t*>(smem_raw + off_range);
    uint16_t* s_tag = reinterpret_cast<uint16_t*>(smem_raw + off_tag);
    uint8_t* s_angle = smem_raw + off_angle;
    const int NWD = (H + 31) >> 5;  // 32-row words per column
    unsigned* s_m32 = reinterpret_cast<unsigned*>(smem_raw + off_masks);  // [MK_COUNT][NWD][CW]
    uint16_t* s_par = s_tag;  // [CW][H] union-find parents of the vertical runs (the tag slab is dead after the projection)
    unsigned* s_rowmask = s_m32 + static_cast<size_t>(MK_COUNT) * NWD * CW;  // [H][NG]
    // scratch of the labelling stage in the angle slab (dead once the links are computed): link words to unite and their scan
    unsigned* s_todo = reinterpret_cast<unsigned*>(s_angle);                 // [(CW - 1) * NWD]
    uint16_t* s_excl = reinterpret_cast<uint16_t*>(s_todo + static_cast<size_t>(CW) * NWD);  // [(CW - 1) * NWD + 1]
    uint16_t* s_qcw = s_excl + static_cast<size_t>(CW) * NWD + 2;                            // [(CW - 1) * NWD] column << 2 | word
    uint16_t* s_task = s_qcw + static_cast<size_t>(CW) * NWD + 2 + 2 * ((H + 7) & ~7);       // after s_broot [2][Hp] (boundary stage): the link list of the union stage, up to the mask area
    unsigned* s_grp = s_rowmask + H * NG;                                                                  // [H][NG]
    uint16_t* s_mylist = reinterpret_cast<uint16_t*>(s_angle + MAX_BLOCKS);                                // [MAX_BLOCKS]
    unsigned char* misc = smem_raw + off_misc;
    const int Hp = (H + 7) & ~7;
    unsigned long long* s_halo_masks = reinterpret_cast<unsigned long long*>(misc);     // [7][2]
    unsigned* s_rowcnt_all = reinterpret_cast<unsigned*>(s_halo_masks + 14);            // [MAX_CL][Hp]
    unsigned* s_rowbase = s_rowcnt_all + MAX_CL * Hp;                                   // [Hp]
    uint8_t* s_halo_angle = reinterpret_cast<uint8_t*>(s_rowbase + Hp);                 // [Hp]
    unsigned* s_flags = reinterpret_cast<unsigned*>(s_halo_angle + Hp);                 // [MAX_CL]
    unsigned* s_misc = s_flags + 2 * MAX_CL;                                            // [0] nmine, [1] scratch   (s_flags: [2 parities][MAX_CL], written by the other CTAs)
    unsigned* s_halo32 = s_misc + 8;                                                    // [5][4] halo mask words
    unsigned* s_scan = s_halo32 + 20;                                                   // [64] block scan scratch
    uint8_t* s_haloL_angle = reinterpret_cast<uint8_t*>(s_scan + 64 + 16);              // [Hp] last angle column of the left neighbour (after the 16 statistics words)
    unsigned* s_haloL_valid = reinterpret_cast<unsigned*>(s_haloL_angle + Hp);          // [4] its VALID words

    // lookup tables: shared-memory copies made once per CTA (the persistent kernel reuses them for every frame)
    double* s_cos_az = reinterpret_cast<double*>(smem_raw + off_lut);  // [CW]
    double* s_sin_az = s_cos_az + CW;                                    // [CW]
    double* s_cos_alt = s_sin_az + CW;                                   // [H]
    double* s_sin_alt = s_cos_alt + H;
    double* s_sin_y_d = s_sin_alt + H;
    double* s_cos_y_d = s_sin_y_d + H;
    float* s_sin_y = reinterpret_cast<float*>(s_cos_y_d + H);
    float* s_cos_y = s_sin_y + H;
    for (int c = tid; c < myCW; c += NT) {
        s_cos_az[c] = A.lut.cos_az[col0 + c];
        s_sin_az[c] = A.lut.sin_az[col0 + c];
    }
    for (int r = tid; r < H; r += NT) {
        s_cos_alt[r] = A.lut.cos_alt[r];
        s_sin_alt[r] = A.lut.sin_alt[r];
        s_sin_y_d[r] = A.lut.sin_y_d[r];
        s_cos_y_d[r] = A.lut.cos_y_d[r];
        s_sin_y[r] = A.lut.sin_y[r];
        s_cos_y[r] = A.lut.cos_y[r];
    }
    if (tid == 0) s_misc[0] = 0;  // blocks dealt to me (reset at the end of every frame)
    __syncthreads();

    uint2* rlist = A.rlist + (static_cast<size_t>(slot) * CL + rank) * A.rcap;
    unsigned* rcount = A.rcount + static_cast<size_t>(slot) * CL;

    // statistics runs (tests): counters live in shared memory, not in registers of the hot loops
    enum { ST_DROPPED = 0, ST_NEAR, ST_NEAR_ANGLE, ST_REMOTE, ST_FALLBACK, ST_MISMATCH, ST_COUNT };
    unsigned* s_stat = s_scan + 64;  // [ST_COUNT] in the slack of the misc area
    if (tid < ST_COUNT) s_stat[tid] = 0;
#define GS_STAT(which) atomicAdd(&s_stat[which], 1u)
    // phase timing (GS_OPT_COLLECT_STATS): thread 0 of every CTA accumulates clock64() deltas per phase
    long long t_mark = 0;
    int t_phase = 0;
#define GS_PHASE(next)                                                        \
    do {                                                                      \
        if ((A.stats & 2) && tid == 0) {                                      \
            const long long now_ = clock64();                                 \
            atomicAdd(&A.cnt->phase[t_phase], static_cast<unsigned long long>(now_ - t_mark)); \
            t_mark = now_;                                                    \
            t_phase = (next);                                                 \
        }                                                                     \
    } while (0)
    if ((A.stats & 2) && tid == 0) t_mark = clock64();

    // P0 — dealing: blocks of 256 consecutive points go to the CTA that most of 4 sampled points fall into (the 4 lanes of a
    // quad take one sample each).  The list of a frame is built while the previous frame's output streams out.
    auto deal_load = [&](const float4* ptsN, int Nn, int b0, float2 (&xy)[2]) {
        const int nbn = (Nn + BLK - 1) / BLK;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = b0 + h * (NT / 4) + (tid >> 2);
            xy[h] = make_float2(1.0f, 1.0f);
            if (b < nbn) xy[h] = __ldg(reinterpret_cast<const float2*>(ptsN + min(b * BLK + 32 + 64 * (tid & 3), Nn - 1)));
        }
    };
    auto deal_vote = [&](int Nn, int b0, const float2 (&xy)[2]) {
        const int nbn = (Nn + BLK - 1) / BLK;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = b0 + h * (NT / 4) + (tid >> 2);
            unsigned o = 0xffu;  // no block: no vote
            if (b < nbn) o = static_cast<unsigned>(min(static_cast<int>(__umulhi(static_cast<unsigned>(approx_col(A, xy[h].x, xy[h].y)), A.cw_magic)), CL - 1));
            // votes of my owner among the quad; most votes wins, ties go to the lowest rank
            unsigned cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) cnt += __shfl_sync(0xffffffffu, o, (lane & ~3) + k) == o;
            unsigned key = cnt * 16u + (15u - (o & 15u));
            key = max(key, __shfl_xor_sync(0xffffffffu, key, 1));
            key = max(key, __shfl_xor_sync(0xffffffffu, key, 2));
            // append my blocks (any order: the pixel keys carry the point order); one shared-memory atomic per warp
            const bool mine = b < nbn && (tid & 3) == 0 && static_cast<int>(15u - (key & 15u)) == rank;
            const unsigned mm = __ballot_sync(0xffffffffu, mine);
            if (mm) {
                unsigned base = 0;
                if (lane == __ffs(mm) - 1) base = atomicAdd(&s_misc[0], static_cast<unsigned>(__popc(mm)));
                base = __shfl_sync(0xffffffffu, base, __ffs(mm) - 1);
                if (mine) s_mylist[base + __popc(mm & ((1u << lane) - 1u))] = static_cast<uint16_t>(b);
            }
        }
    };
    // the rest of the blocks when a frame has more than NT / 2 of them (> 131 k points with 1024 threads)
    auto deal_rest = [&](const float4* ptsN, int Nn) {
        const int nbn = (Nn + BLK - 1) / BLK;
        for (int b0 = NT / 2; b0 < nbn; b0 += NT / 2) {  // uniform trip count
            float2 xy[2];
            deal_load(ptsN, Nn, b0, xy);
            deal_vote(Nn, b0, xy);
        }
    };
    long long beg_cur = 0;
    int N_cur = 0;
    const int f_first = A.first_frame + cluster_id;
    if (f_first < A.frames) {
        beg_cur = A.offsets[f_first];
        N_cur = static_cast<int>(A.offsets[f_first + 1] - beg_cur);
        float2 xy[2];
        deal_load(A.points + beg_cur, N_cur, 0, xy);
        deal_vote(N_cur, 0, xy);
        deal_rest(A.points + beg_cur, N_cur);
    }

    int f_next = A.frames;
    for (int f = f_first; f < A.frames; f = f_next) {
        const long long beg = beg_cur;  // A.offsets[f]: loaded for the dealing during the previous frame (or the prologue)
        const int N = N_cur;
        const float4* pts = A.points + beg;

        // ================================================================================ P0: clear (the block list was built during the previous frame's output)
        {
            uint4* r4 = reinterpret_cast<uint4*>(s_range);
            const int n16 = (H * S16 * 2 + 15) / 16;
            for (int i = tid; i < n16; i += NT) r4[i] = make_uint4(0, 0, 0, 0);
            uint4* t4 = reinterpret_cast<uint4*>(s_tag);
            for (int i = tid; i < n16; i += NT) t4[i] = make_uint4(0, 0, 0, 0);
        }
        if (tid == 0) rcount[rank] = 0;
        __syncthreads();
        const int nmine = static_cast<int>(s_misc[0]);

        GS_PHASE(1);
        // ================================================================================ P1: projection + pixel ownership
        // "The last point in input order wins" (alfa_gs.cpp:291) as ONE 32-bit shared-memory atomicMax per point on the
        // key (g << 16 | range), g = (index >> tag_shift) + 1: a later point always beats an earlier one and carries its
        // range along; points with the same tag sit in adjacent lanes and are resolved with shuffles.  No barriers.
        unsigned* s_key = reinterpret_cast<unsigned*>(s_range);  // [H][S16] u32 over the range + tag slabs, compacted to u16 below
        int tag_shift = 0;  // order tag = index >> tag_shift, the smallest shift that fits 16 bits (uniform per frame)
        while ((((N - 1) >> tag_shift) + 1) > 65535) ++tag_shift;
        {
            // Work units of 64 consecutive points (a quarter of a dealt block) go round-robin to the warps; the two loads of the
            // next unit are issued before the current unit is processed.
            const bool stats1 = (A.stats & 1) != 0;
            const int n_punits = nmine * (BLK / 64);
            // x, y, z are loaded as 8 + 4 bytes, not as one 16-byte record: the register of the unused intensity word would be
            // handed out as a temporary while the prefetch is in flight, and the first write to it waits for the load (WAW)
            const float4* pts_r;  // the frame's base address as a register pair (one IMAD.WIDE per point address)
            asm volatile("mov.u64 %0, %1;" : "=l"(pts_r) : "l"(pts));
            const unsigned a_mylist = smem_base32 + off_angle + MAX_BLOCKS;  // s_mylist through the shared window (no SR_CgaCtaId per access)
            auto load_point = [&](int i) -> float4 {
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < N) {
                    const float2 xy = __ldcs(reinterpret_cast<const float2*>(pts_r + i));
                    q.x = xy.x;
                    q.y = xy.y;
                    q.z = __ldcs(reinterpret_cast<const float*>(pts_r + i) + 2);
                }
                return q;
            };
            auto load_unit = [&](int u, float4& pa, float4& pb) -> int {
                const int i0 = static_cast<int>(lds_u16(a_mylist + 2 * (u >> 2))) * BLK + (u & 3) * 64 + lane;
                pa = load_point(i0);
                pb = load_point(i0 + 32);
                return i0;
            };
            // `checked`: statistics runs and configurations without the certified fast path evaluate the exact projection for
            // every point; the production instantiation keeps the hot path to the certified pixel, the key and one atomic,
            // with everything else (padding lanes, uncertified points, the alias column) in one cold branch
            const unsigned a_key = smem_base32 + off_range;  // shared-window address of the key slab
            auto place = [&](auto checked_tag, auto nearest_tag, auto ts_tag, const int i, const Lean& L) {
                constexpr bool CHECKED = decltype(checked_tag)::value;
                // pixel_policy: the reference keeps the LAST point in input order (key = order tag << 16 | range, the later
                // point has the larger key); the north-star variant keeps the NEAREST return, ties to the lowest index (key =
                // (65536 - range) << 16 | ~order tag: the smaller range, then the earlier point, has the larger key).  Both are one atomicMax.
                constexpr bool NEAREST = decltype(nearest_tag)::value;
                unsigned k = 0x80000000u | lane;  // owner << 16 | slab address; this value is unique per lane: never matches
                unsigned val = 0;                 // g << 16 | encoded range
                int row = L.row, col = L.col;
                unsigned enc = L.enc;
                bool have = L.ok && i < N;
                if (CHECKED || !have) {
                    if (i < N) {
                        // not certified: a gate drop, or the exact evaluation (every point when statistics are collected).  The
                        // coordinates are read again here (a cache hit, ~0.5 % of the points) instead of being kept in registers
                        // through the hot path: the loop sits at its 64 registers, and what the compiler spills otherwise is a
                        // register of the prefetched next unit — a store that waits for the load it was meant to overlap
                        const float4 p4 = load_point(i);
                        int st = A.use_fast ? (L.ok ? 0 : project_undecided(A.gates_dev, p4.x, p4.y, p4.z)) : 2;
                        if (st == 2 || (CHECKED && stats1)) {
                            const Proj pr = project_point_d(A.kp_dev, p4.x, p4.y, p4.z);
                            if (CHECKED && stats1) {
                                if (st == 2) GS_STAT(ST_FALLBACK);
                                else if ((st == 1) != (pr.row < 0) || (st == 0 && (row != pr.row || col != pr.col || enc != pr.enc))) GS_STAT(ST_MISMATCH);
                                if (pr.row < 0) GS_STAT(ST_DROPPED);
                                else if (pr.near_edge) GS_STAT(ST_NEAR);
                            }
                            if (st == 2) {
                                row = pr.row;
                                col = pr.col;
                                enc = pr.enc;
                                st = pr.row < 0 ? 1 : 0;
                            }
                        }
                        have = st == 0;
                        if (have && col == W) {  // Q2: at<ushort>(row, W) is flat index (row + 1) * W (only the exact path returns it)
                            row += 1;
                            col = 0;
                        }
                    }
                }
                if (have) {
                    const int owner = SPEC ? col / CW : min(static_cast<int>(__umulhi(static_cast<unsigned>(col), A.cw_magic)), CL - 1);
                    const int lc = col - owner * CW;
                    k = (static_cast<unsigned>(owner) << 16) | static_cast<unsigned>(row * S16 + lc);
                    const unsigned tagv = static_cast<unsigned>(i >> tag_shift) + 1u;
                    val = NEAREST ? (((65536u - enc) << 16) | (65535u - tagv)) : ((tagv << 16) | enc);  // 65536 - enc is in [1, 65535]: never 0 = empty
                }
                // points that share an order tag are 2^tag_shift consecutive lanes: the highest lane among those that hit the
                // same pixel holds the latest point (down-shuffles over the group; __match_any_sync would do the same for a
                // whole warp but costs ~100 cycles of the address-divergence unit per call when all keys differ)
                // TS: 0 = every point has its own tag (no dedupe), 1 = pairs of lanes share a tag (frames of 64 Ki .. 128 Ki points),
                // 2 = the run-time general case
                constexpr int TS = decltype(ts_tag)::value;
                bool lose = false;
                if constexpr (NEAREST) {
                    if (tag_shift) {  // uniform
                        // same pixel, same order tag: the larger key (smaller range) wins, equal ranges go to the lower lane (earlier point)
#pragma unroll 1
                        for (int d = 1; d < (1 << tag_shift); ++d) {
                            const unsigned other = __shfl_xor_sync(0xffffffffu, k, d);
                            const unsigned oval = __shfl_xor_sync(0xffffffffu, val, d);
                            if (other == k && (oval > val || (oval == val && (lane ^ d) < lane))) lose = true;
                        }
                    }
                } else if constexpr (TS == 1) {
                    // a lane loses when a HIGHER lane of the warp hits the same pixel: inside its tag group that is the rule, beyond
                    // the group the higher lane carries a larger tag and would win the atomicMax anyway
                    lose = shfl_down_eq(k, 1);
                } else if constexpr (TS == 2) {
                    if (tag_shift) {  // uniform
#pragma unroll 1
                        for (int d = 1; d < (1 << tag_shift); ++d) lose |= shfl_down_eq(k, d);
                    }
                }
                const bool winner = have && !lose;
                const bool local = winner && static_cast<int>(k >> 16) == rank;
                if (local) asm volatile("red.shared::cta.max.u32 [%0], %1;" ::"r"(a_key + 4u * (k & 0xffffu)), "r"(val) : "memory");
                // remote pixels: queue the record (the compiler aggregates the slot claim over the lanes of the branch)
                if (winner && !local) {
                    const unsigned slot = atomicAdd(&rcount[rank], 1u);
                    if (slot < A.rcap) rlist[slot] = make_uint2(k, val);
                    if (CHECKED && stats1) GS_STAT(ST_REMOTE);
                }
            };
            auto rounds = [&](auto checked_tag, auto nearest_tag, auto ts_tag) {
                float4 na = make_float4(0.f, 0.f, 0.f, 0.f), nb4 = na;
                int ni = 0;
                if (warp < n_punits) ni = load_unit(warp, na, nb4);
                for (int u = warp; u < n_punits; u += NT / 32) {
                    const float4 pa = na, pb = nb4;
                    const int i0 = ni;
                    if (u + NT / 32 < n_punits) ni = load_unit(u + NT / 32, na, nb4);
                    Lean La, Lb;
                    project_lean2(A, pa, pb, La, Lb);  // straight-line, the float chains of the two points packed (f32x2)
                    place(checked_tag, nearest_tag, ts_tag, i0, La);
                    place(checked_tag, nearest_tag, ts_tag, i0 + 32, Lb);
                }
            };
            using ts_any = std::integral_constant<int, 2>;
            if (kp.pixel_policy == GS_PIXEL_NEAREST) {  // uniform
                if (stats1 || !A.use_fast) rounds(std::true_type{}, std::true_type{}, ts_any{});
                else rounds(std::false_type{}, std::true_type{}, ts_any{});
            } else if (stats1 || !A.use_fast) {
                rounds(std::true_type{}, std::false_type{}, ts_any{});
            } else if (tag_shift == 0) {  // the production instantiations: the dedupe of a frame's tag shift is compiled in
                rounds(std::false_type{}, std::false_type{}, std::integral_constant<int, 0>{});
            } else if (tag_shift == 1) {
                rounds(std::false_type{}, std::false_type{}, std::integral_constant<int, 1>{});
#if GS_EXP_DOUBLE == 1
                rounds(std::false_type{}, std::false_type{}, std::integral_constant<int, 1>{});
#endif
            } else {
                rounds(std::false_type{}, std::false_type{}, ts_any{});
            }
        }
        GS_PHASE(2);
        // remote records and counts: barrier.cluster arrives with release and waits with acquire semantics at cluster scope,
        // which orders the global stores / atomics above before the reads below in the other CTAs of the cluster — no __threadfence
        csync();   // #1: all local pixels committed everywhere
        if (tid == 0) s_misc[0] = 0;  // every thread has read the block count; the next frame's list is built during the output

        GS_PHASE(3);
        // ================================================================================ P2: remote records
        {
            unsigned n_any = 0;
            for (int k = 0; k < CL; ++k) n_any |= __ldcg(&rcount[k]);
            if (n_any) {
                if constexpr (SOFT) {
                    // nobody can reach into my shared memory: I apply the records addressed to me from every CTA's list
                    for (int k = 0; k < CL; ++k) {
                        const uint2* lst = A.rlist + (static_cast<size_t>(slot) * CL + k) * A.rcap;
                        const unsigned n = min(__ldcg(&rcount[k]), A.rcap);
                        for (unsigned i = tid; i < n; i += NT) {
                            const uint2 rec = __ldcg(&lst[i]);
                            if (static_cast<int>(rec.x >> 16) == rank) atomicMax(s_key + (rec.x & 0xffffu), rec.y);
                        }
                    }
                } else {
                    const unsigned n = min(__ldcg(&rcount[rank]), A.rcap);
                    for (unsigned i = tid; i < n; i += NT) {
                        const uint2 rec = __ldcg(&rlist[i]);
                        atomicMax(cluster.map_shared_rank(s_key, rec.x >> 16) + (rec.x & 0xffffu), rec.y);
                    }
                }
                csync();  // #2: every record is applied before its list is refilled / before the keys are compacted
            }
        }
        // compact the keys to the u16 range slab in place (the low half of every key is the winner's range); the freed
        // upper part (the tag slab) becomes the union-find forest later
        {
            const int n_quads = (H * S16) >> 2;  // H is a power of two >= 16 and S16 is even: the slab is a whole number of quads
            const uint4* k4 = reinterpret_cast<const uint4*>(s_key);
            uint2* r2 = reinterpret_cast<uint2*>(s_range);
            constexpr int CQ = 3;  // quads per thread between two barriers (12 registers)
            const bool nearest_keys = kp.pixel_policy == GS_PIXEL_NEAREST;
            for (int base = 0; base < n_quads; base += CQ * NT) {
                uint4 v[CQ];
#pragma unroll
                for (int j = 0; j < CQ; ++j) {
                    const int i = base + j * NT + tid;
                    v[j] = make_uint4(0u, 0u, 0u, 0u);
                    if (i < n_quads) v[j] = k4[i];
                }
                __syncthreads();  // every key of this chunk is read before its bytes are overwritten (writes land in the first half of what was read so far)
#pragma unroll
                for (int j = 0; j < CQ; ++j) {
                    const int i = base + j * NT + tid;
                    if (i < n_quads) {
                        if (nearest_keys) {  // the range sits complemented in the upper half of a key; an untouched key is an empty pixel
                            const unsigned e0 = v[j].x ? 65536u - (v[j].x >> 16) : 0u, e1 = v[j].y ? 65536u - (v[j].y >> 16) : 0u;
                            const unsigned e2 = v[j].z ? 65536u - (v[j].z >> 16) : 0u, e3 = v[j].w ? 65536u - (v[j].w >> 16) : 0u;
                            r2[i] = make_uint2(e0 | (e1 << 16), e2 | (e3 << 16));
                        } else {
                            r2[i] = make_uint2((v[j].x & 0xffffu) | (v[j].y << 16), (v[j].z & 0xffffu) | (v[j].w << 16));
                        }
                    }
                }
            }
            __syncthreads();
        }

        // optional output: the range image
        if (A.range_out) {
            uint16_t* out = A.range_out + static_cast<size_t>(f) * kp.P;
            for (int i = tid; i < H * myCW; i += NT) {
                const int r = i / myCW, c = i - r * myCW;
                out[r * W + col0 + c] = s_range[r * S16 + c];
            }
        }

        // Cloud2RangeNode::CloudCb runs RepairGaps(range_image, 5, 1.0f) and builds the angle image from the repaired image, while
        // EraseGroundBFS keeps the original one (cloud2range_node.cpp:468-473).  Typed reading: depths are the encoded
        // u16 values as floats, in-painted top to bottom inside a column (repaired pixels feed later ones), written back truncated.
        // The repaired slab lives in the tag slab, idle between the key compaction and the forest.
        const uint16_t* s_walk = s_range;
        if (A.repair) {  // uniform
            uint16_t* s_rep = s_tag;
            {
                const uint4* src = reinterpret_cast<const uint4*>(s_range);
                uint4* dst = reinterpret_cast<uint4*>(s_rep);
                const int n16 = (H * S16 * 2 + 15) / 16;
                for (int i = tid; i < n16; i += NT) dst[i] = src[i];
            }
            __syncthreads();
            // One thread per column: it first collects the rows of its empty pixels (a 128-bit mask), then visits only those, top
            // to bottom (repaired pixels feed later ones, alfa_gs.cpp:304-305: rows above are read in their current state, rows below
            // are still original when a pixel is visited).  A pair (prev, next) counts when both are non-empty and
            // |prev - next| < 1.0f, i.e. when the two encoded values are EQUAL; sums of at most 32 u16 values are exact in float, so
            // the reference's float accumulation is the integer sum, divided and truncated as there (:326-328).  Warps reconverge
            // after every visit (a lane subset left alone in a data-dependent loop re-issues the whole loop on its own).
            for (int c0w = 0; c0w < myCW; c0w += NT) {  // uniform trip count
                const int c = c0w + tid;
                const bool col_ok = c < myCW;
                uint16_t* col = s_rep + (col_ok ? c : 0);
                unsigned long long e_lo = 0, e_hi = 0;
                for (int r = 0; r < min(H, 64); ++r)
                    if (col[r * S16] == 0) e_lo |= 1ull << r;
                for (int r = 64; r < H; ++r)
                    if (col[r * S16] == 0) e_hi |= 1ull << (r - 64);
                if (!col_ok) e_lo = e_hi = 0;
                while (__any_sync(0xffffffffu, (e_lo | e_hi) != 0ull)) {
                    if ((e_lo | e_hi) != 0ull) {
                        int r;
                        if (e_lo) {
                            r = __ffsll(static_cast<long long>(e_lo)) - 1;
                            e_lo &= e_lo - 1ull;
                        } else {
                            r = 63 + __ffsll(static_cast<long long>(e_hi));
                            e_hi &= e_hi - 1ull;
                        }
                        unsigned av[4], bv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            av[i] = (r - 1 - i >= 0) ? col[(r - 1 - i) * S16] : 0u;
                            bv[i] = (r + 1 + i < H) ? col[(r + 1 + i) * S16] : 0u;
                        }
                        unsigned cnt = 0, sum = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const bool m = av[i] == bv[j] && av[i] != 0u;
                                cnt += m ? 1u : 0u;
                                sum += m ? av[i] : 0u;
                            }
                        if (cnt) col[r * S16] = static_cast<uint16_t>(__float2int_rz(__fdiv_rn(static_cast<float>(2u * sum), static_cast<float>(2u * cnt))));
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            s_walk = s_rep;
        }

        GS_PHASE(4);
        // ================================================================================ A: angle image (column walks)
        // one thread per (column, 32-row segment); bottom -> top like the reference (alfa_gs.cpp:347-382).  `checked`: statistics
        // runs and configurations without the certified fast path; the production instantiation carries neither test.
        auto angle_walk = [&](auto checked_tag) {
            constexpr bool CHECKED = decltype(checked_tag)::value;
            const int nword = (H + 31) / 32;
            const int nseg = myCW * nword;
            for (int s = tid; s < nseg; s += NT) {
                const int c = s % myCW, wseg = s / myCW;
                const int lo = wseg * 32, hi = min(lo + 31, H - 1);
                // previous non-empty pixel below the segment
                float p_z = 0.0f, p_x = 0.0f;  // prev_r * sinf(prev_angle_y), prev_r * cosf(prev_angle_y) (float products, :366-367)
                bool has_prev = false;
                for (int r = hi + 1; r < H; ++r) {
                    const unsigned e = s_walk[r * S16 + c];
                    if (e) {
                        const float pr = __double2float_rn(decode_range_d(kp, e));  // prev_r = range (:379)
                        p_z = __fmul_rn(pr, s_sin_y[r]);
                        p_x = __fmul_rn(pr, s_cos_y[r]);
                        has_prev = true;
                        break;
                    }
                }
                if (!has_prev && col0 + c == 1) {
                    // column-1 quirk (SURVEY.md A.2 iii): prev_col starts at 1 (:340) with prev_r = 0, prev_angle_y = 0
                    bool col0_empty = true;
                    for (int r = 0; r < H; ++r)
                        if (s_walk[r * S16]) { col0_empty = false; break; }
                    if (col0_empty) {
                        has_prev = true;
                        p_z = __fmul_rn(0.0f, 0.0f);  // 0 * sinf(0)
                        p_x = __fmul_rn(0.0f, 1.0f);  // 0 * cosf(0)
                    }
                }
                unsigned vbits = 0;  // shifted in row by row: the first row walked (hi) ends up at bit hi - lo
                const uint16_t* p_rng = s_walk + hi * S16 + c;
                uint8_t* p_ang = s_angle + hi * S8 + c;
                // the per-row tables through shared-window addresses that walk along with the row (no address arithmetic per pixel)
                unsigned a_f = smem_base32 + off_lut + static_cast<unsigned>(reinterpret_cast<unsigned char*>(s_sin_y) - reinterpret_cast<unsigned char*>(s_cos_az)) + 4u * hi;
                unsigned a_d = smem_base32 + off_lut + static_cast<unsigned>(reinterpret_cast<unsigned char*>(s_sin_y_d) - reinterpret_cast<unsigned char*>(s_cos_az)) + 8u * hi;
                // Two rows per iteration: the differences (dz, dx) of a row depend on the previous non-empty pixel only through two
                // products, so both rows' float chains (quotient, polynomial, scaling, rounding trick) run packed as f32x2.
                // fast: float atan with a 2e-7 rad error budget; results within 4e-6 rad of a centi-radian edge (and 0/0) are
                // recomputed with the bit-exact clone of the host's atan2f
                // branch-free: an empty pixel (8 % of them) is evaluated like any other and its results are dropped by selects
                auto row_in = [&](unsigned e, unsigned af, unsigned ad, bool& live, float& dz, float& dx) {
                    const bool valid = e != 0u;
                    const double rng = decode_range_d(kp, e);
                    const float sy = lds_f32(af), cy = lds_f32(af + 4u * H);  // s_sin_y[r], s_cos_y[r]
                    dz = __double2float_rn(fabs(__dsub_rn(static_cast<double>(p_z), __dmul_rn(rng, lds_f64(ad)))));           // :366  s_sin_y_d[r]
                    dx = __double2float_rn(fabs(__dsub_rn(static_cast<double>(p_x), __dmul_rn(rng, lds_f64(ad + 8u * H)))));  // :367  s_cos_y_d[r]
                    live = valid && has_prev;
                    const float rf = __double2float_rn(rng);  // prev_r = range (:379)
                    p_z = valid ? __fmul_rn(rf, sy) : p_z;
                    p_x = valid ? __fmul_rn(rf, cy) : p_x;
                    has_prev = has_prev || valid;
                };
                auto row_out = [&](bool live, float dz, float dx, float tm, float diff) -> unsigned {
                    // bits(sc + M) - bits(M) = rn(sc), M = 1.5 * 2^23; diff = sc - rn(sc), exact
                    const int kk = (__float_as_int(tm) - 0x4B400000) - (diff < 0.0f ? 1 : 0);
                    const bool sure = (!CHECKED || A.use_fast) && fabsf(diff) > 4e-4f;  // false for NaN
                    unsigned a_out = live ? static_cast<unsigned>(kk) : 0u;
                    if (live && (!sure || (CHECKED && (A.stats & 1)))) {
                        const float alfa = atan2f_exact(dz, dx);       // :368
                        const float scaled = __fmul_rn(alfa, 100.0f);  // :370
                        const unsigned exact = static_cast<unsigned>(__float2int_rz(scaled));
                        if ((CHECKED && (A.stats & 1))) {
                            const float fe = scaled - floorf(scaled);
                            if (fminf(fe, 1.0f - fe) < 1e-3f) GS_STAT(ST_NEAR_ANGLE);
                            if (!sure) GS_STAT(ST_FALLBACK);
                            else if (exact != a_out) GS_STAT(ST_MISMATCH);
                        }
                        if (!sure) a_out = exact;
                    }
                    return a_out;
                };
#pragma unroll 2
                for (int r = hi; r >= lo; r -= 2, p_rng -= 2 * S16, p_ang -= 2 * S8, a_f -= 8u, a_d -= 16u) {
                    // (a segment has 32 rows, or H rows for H = 2 .. 16: an even number — H is a power of two, checked on the host)
                    const unsigned e0 = *p_rng;
                    const unsigned e1 = *(p_rng - S16);
                    vbits = (vbits << 2) | (e0 ? 2u : 0u) | (e1 ? 1u : 0u);
                    bool live0, live1;
                    float dz0, dx0, dz1, dx1;
                    row_in(e0, a_f, a_d, live0, dz0, dx0);
                    row_in(e1, a_f - 4u, a_d - 8u, live1, dz1, dx1);
                    const float2 q = make_float2(__fmul_rn(fminf(dz0, dx0), rcp_ftz(fmaxf(dz0, dx0))), __fmul_rn(fminf(dz1, dx1), rcp_ftz(fmaxf(dz1, dx1))));
                    float2 t = atan01x2(q);
                    if (dz0 > dx0) t.x = 1.57079632679489662f - t.x;
                    if (dz1 > dx1) t.y = 1.57079632679489662f - t.y;
                    const float2 sc = mul2(t, splat2(100.0f));  // 0 .. 157.1 (NaN for 0/0 and flushed operands: not sure)
                    const float2 tm = add2(sc, splat2(12582912.0f));
                    const float2 diff = fma2(add2(tm, splat2(-12582912.0f)), splat2(-1.0f), sc);  // sc - (tm - M): the product is exact, one rounding
                    *p_ang = static_cast<uint8_t>(row_out(live0, dz0, dx0, tm.x, diff.x));
                    *(p_ang - S8) = static_cast<uint8_t>(row_out(live1, dz1, dx1, tm.y, diff.y));
                }
                // valid bits of this segment -> the column's VALID mask word
                s_m32[(MK_VALID * NWD + wseg) * CW + c] = vbits;
            }
        };
        if ((A.stats & 1) || !A.use_fast) angle_walk(std::true_type{});
        else angle_walk(std::false_type{});
#if GS_EXP_DOUBLE == 2
        __syncthreads();
        angle_walk(std::false_type{});
#endif
        if (A.repair) {
            // the ground search sees the ORIGINAL range image (cloud2range_node.cpp:473): VALID words from it, not from the repaired one
            const int nword = (H + 31) / 32;
            for (int sgm = tid; sgm < myCW * nword; sgm += NT) {
                const int c = sgm % myCW, wseg = sgm / myCW;
                const int lo = wseg * 32, hi = min(lo + 31, H - 1);
                unsigned vbits = 0;
                for (int r = lo; r <= hi; ++r)
                    if (s_range[r * S16 + c]) vbits |= 1u << (r - lo);
                s_m32[(MK_VALID * NWD + wseg) * CW + c] = vbits;
            }
        }
        __syncthreads();
        if (A.sg_ws > 0) {
            // Alfa_GS::SavitskyGolaySmoothing (alfa_gs.cpp:705-713) = cv::filter2D with a ws x 1 kernel and BORDER_REFLECT101:
            // float32 accumulation in tap order, round half to even, saturate (pinned against OpenCV in tests/golden).
            // Smoothed angles stay below 256 (157 x the positive tap sum 1.19).  Filter into the (idle) tag slab, copy back.
            uint8_t* tmp = reinterpret_cast<uint8_t*>(s_tag);
            // a thread filters 8 rows of one column from a window held in registers (8 + ws - 1 loads for 8 outputs)
            auto sg_pass = [&](auto ws_tag) {
                constexpr int WS = decltype(ws_tag)::value, HALF = WS / 2, RB = 8;
                const int n_rb = (H + RB - 1) / RB;
                for (int t = tid; t < myCW * n_rb; t += NT) {
                    const int rb = t / myCW, c = t - rb * myCW;
                    const int r0 = rb * RB;
                    float v[RB + 2 * HALF];
#pragma unroll
                    for (int j = 0; j < RB + 2 * HALF; ++j) {
                        int rr = r0 + j - HALF;
                        while (rr < 0 || rr >= H) rr = rr < 0 ? -rr : 2 * H - 2 - rr;  // BORDER_REFLECT101
                        v[j] = static_cast<float>(s_angle[rr * S8 + c]);
                    }
#pragma unroll
                    for (int i = 0; i < RB; ++i) {
                        if (r0 + i < H) {
                            float acc = 0.0f;
#pragma unroll
                            for (int k = 0; k < WS; ++k) acc = __fadd_rn(acc, __fmul_rn(A.sg_t[k], v[i + k]));
                            tmp[(r0 + i) * S8 + c] = static_cast<uint8_t>(max(0, min(255, __float2int_rn(acc))));
                        }
                    }
                }
            };
            if (H >= 8 && A.sg_ws == 5) sg_pass(std::integral_constant<int, 5>{});
            else if (H >= 8 && A.sg_ws == 7) sg_pass(std::integral_constant<int, 7>{});
            else if (H >= 8 && A.sg_ws == 9) sg_pass(std::integral_constant<int, 9>{});
            else if (H >= 8 && A.sg_ws == 11) sg_pass(std::integral_constant<int, 11>{});
            else {
                const int half = A.sg_ws >> 1;
                for (int i = tid; i < H * myCW; i += NT) {
                    const int r = i / myCW, c = i - r * myCW;
                    float acc = 0.0f;
                    for (int k = 0; k < A.sg_ws; ++k) {
                        int rr = r + k - half;
                        while (rr < 0 || rr >= H) rr = rr < 0 ? -rr : 2 * H - 2 - rr;
                        acc = __fadd_rn(acc, __fmul_rn(A.sg_t[k], static_cast<float>(s_angle[rr * S8 + c])));
                    }
                    const int v = __float2int_rn(acc);
                    tmp[r * S8 + c] = static_cast<uint8_t>(max(0, min(255, v)));
                }
            }
            __syncthreads();
            {
                // copy back whole rows as 32-bit words (S8 is a multiple of 4; the padding columns are never read)
                unsigned* dst = reinterpret_cast<unsigned*>(s_angle);
                const unsigned* src = reinterpret_cast<const unsigned*>(tmp);
                for (int i = tid; i < (H * S8) >> 2; i += NT) dst[i] = src[i];
            }
            __syncthreads();
        }
        if (A.angle_out) {
            uint16_t* out = A.angle_out + static_cast<size_t>(f) * kp.P;
            for (int i = tid; i < H * myCW; i += NT) {
                const int r = i / myCW, c = i - r * myCW;
                out[r * W + col0 + c] = s_angle[r * S8 + c];
            }
        }
        const int right_rank = (rank + 1 == CL) ? 0 : rank + 1;
        const int left_rank = (rank == 0) ? CL - 1 : rank - 1;
#define MW(kind, w, c) s_m32[((kind) * NWD + (w)) * CW + (c)]
        // Halos are PUSHED into the neighbours' shared memory before the cluster barrier: my column 0 (angles, VALID words) is
        // the left neighbour's right halo, my last column its right neighbour's left halo (column W-1's right neighbour is
        // column 0: Label::WrapCols).  After the barrier nobody reads another CTA's angle slab or link words any more, which
        // saves the cluster barrier that used to follow the link stage.
        if constexpr (SOFT) {
            for (int r = tid; r < H; r += NT) {
                mail->angle0[rank][r] = s_angle[r * S8];
                mail->angleL[rank][r] = s_angle[r * S8 + myCW - 1];
            }
            if (tid < NWD) {
                mail->valid0[rank][tid] = MW(MK_VALID, tid, 0);
                mail->validL[rank][tid] = MW(MK_VALID, tid, myCW - 1);
            }
        } else {
            uint8_t* to_left = cluster.map_shared_rank(s_halo_angle, left_rank);
            uint8_t* to_right = cluster.map_shared_rank(s_haloL_angle, right_rank);
            for (int r = tid; r < H; r += NT) {
                to_left[r] = s_angle[r * S8];
                to_right[r] = s_angle[r * S8 + myCW - 1];
            }
            if (tid < NWD) {
                cluster.map_shared_rank(s_halo32, left_rank)[0 * 4 + tid] = MW(MK_VALID, tid, 0);
                cluster.map_shared_rank(s_haloL_valid, right_rank)[tid] = MW(MK_VALID, tid, myCW - 1);
            }
        }
        GS_PHASE(5);
        csync();  // #4: angle slabs, VALID masks and halos complete in the whole cluster
        if constexpr (SOFT) {
            // pull my halos out of the neighbours' mailbox slots (L2 reads: the lines were written by other SMs)
            for (int r = tid; r < H; r += NT) {
                s_halo_angle[r] = __ldcg(&mail->angle0[right_rank][r]);
                s_haloL_angle[r] = __ldcg(&mail->angleL[left_rank][r]);
            }
            if (tid < NWD) {
                s_halo32[0 * 4 + tid] = __ldcg(&mail->valid0[right_rank][tid]);
                s_haloL_valid[tid] = __ldcg(&mail->validL[left_rank][tid]);
            }
            __syncthreads();
        }

        GS_PHASE(6);
        // ================================================================================ G: links, runs, seeds, ground label
        const int n_units = myCW * NWD;  // one unit = 32 rows of one column
        // Angle-difference links (label.cpp:90-99 with the float threshold folded to an integer): bit i of
        // VA / HA says |a(r) - a(r+1)| <= K  /  |a(r, c) - a(r, c+1)| <= K for r = 32w + i.  Byte-parallel when the slab is a
        // whole number of 4-column words: a thread takes 4 columns x 8 rows, one 32-bit load per row gives the 4 angles,
        // VABSDIFF4 the four differences, and "d <= K" is bit 7 of ((K | 0x80) - (d & 0x7f)) & ~d per byte (K < 128);
        // eight rows of those bits shifted together are the bytes of the four columns' mask words.
        const bool simd_links = (myCW & 3) == 0 && (H & 7) == 0 && kp.k_ground < 128;  // uniform
        GS_REPEAT(5) {
        if (simd_links) {
            const int n_c4 = myCW >> 2;
            const unsigned k4 = (static_cast<unsigned>(kp.k_ground) * 0x01010101u) | 0x80808080u;
            uint8_t* m8 = reinterpret_cast<uint8_t*>(s_m32);
            for (int t = tid; t < n_c4 * (H >> 3); t += NT) {
                const int rb = t / n_c4, c4 = (t - rb * n_c4) << 2;
                const int r0 = rb << 3;
                const bool last_group = c4 + 4 == myCW;
                unsigned acc_v = 0, acc_h = 0;
                unsigned cur = *reinterpret_cast<const unsigned*>(s_angle + r0 * S8 + c4);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = r0 + i;
                    const unsigned right = last_group ? static_cast<unsigned>(s_halo_angle[r]) : *reinterpret_cast<const unsigned*>(s_angle + r * S8 + c4 + 4);
                    const unsigned dh = __vabsdiffu4(cur, __funnelshift_r(cur, right, 8));
                    acc_h = (acc_h >> 1) | ((k4 - (dh & 0x7f7f7f7fu)) & ~dh & 0x80808080u);
                    unsigned le_v = 0;
                    if (r + 1 < H) {
                        const unsigned down = *reinterpret_cast<const unsigned*>(s_angle + (r + 1) * S8 + c4);
                        const unsigned dv = __vabsdiffu4(cur, down);
                        le_v = (k4 - (dv & 0x7f7f7f7fu)) & ~dv & 0x80808080u;
                        cur = down;
                    }
                    acc_v = (acc_v >> 1) | le_v;
                }
                const int w = r0 >> 5, byte = (r0 >> 3) & 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    m8[4 * ((MK_VA * NWD + w) * CW + c4 + j) + byte] = static_cast<uint8_t>(acc_v >> (8 * j));
                    m8[4 * ((MK_HA * NWD + w) * CW + c4 + j) + byte] = static_cast<uint8_t>(acc_h >> (8 * j));
                }
            }
            __syncthreads();
        }
        for (int u = tid; u < n_units; u += NT) {
            const int w = u / myCW, c = u - w * myCW;
            const int lo = w * 32, nrow = min(32, H - lo);
            const unsigned valid = MW(MK_VALID, w, c);
            const bool last = (c == myCW - 1);
            unsigned va = 0, ha = 0;
            const int a_first = s_angle[lo * S8 + c];
            if (simd_links) {
                const unsigned rows_mask = nrow < 32 ? (1u << nrow) - 1u : 0xffffffffu;  // bytes above H were not written
                va = MW(MK_VA, w, c) & rows_mask;
                ha = MW(MK_HA, w, c) & rows_mask;
            } else {
                int a_here = a_first;
                for (int i = 0; i < nrow; ++i) {
                    const int r = lo + i;
                    const int a_down = (r + 1 < H) ? s_angle[(r + 1) * S8 + c] : 0;
                    const int a_right = last ? s_halo_angle[r] : s_angle[r * S8 + c + 1];
                    if (r + 1 < H && abs(a_here - a_down) <= kp.k_ground) va |= 1u << i;
                    if (abs(a_here - a_right) <= kp.k_ground) ha |= 1u << i;
                    a_here = a_down;
                }
            }
            const unsigned valid_up = (w + 1 < NWD) ? MW(MK_VALID, w + 1, c) : 0u;
            const unsigned v = va & valid & ((valid >> 1) | (valid_up << 31));  // bit i: rows r and r+1 linked (both non-empty)
            const unsigned valid_right = last ? s_halo32[0 * 4 + w] : MW(MK_VALID, w, c + 1);
            const unsigned hm = ha & valid & valid_right;                       // bit i: (r, c) and (r, c+1) linked
            unsigned linked_prev = v << 1;  // bit i: row r linked to row r-1; bit 0 looks into the word below
            if (w > 0 && (valid & 1u) && (MW(MK_VALID, w - 1, c) >> 31) && abs(a_first - static_cast<int>(s_angle[(lo - 1) * S8 + c])) <= kp.k_ground) linked_prev |= 1u;
            const unsigned runs = valid & ~linked_prev;  // first (lowest-index) pixel of every vertical run
            MW(MK_VA, w, c) = va;
            MW(MK_HA, w, c) = ha;
            MW(MK_V, w, c) = v;
            MW(MK_HM, w, c) = hm;
            MW(MK_S, w, c) = runs;
            for (unsigned m = runs; m; m &= m - 1) {
                const int node = run_node(c, lo + __ffs(m) - 1);
                s_par[node] = static_cast<uint16_t>(node);
            }
            // seed (alfa_gs.cpp:720-736): the bottom-most non-empty pixel of the column, row 0 when the column is empty
            bool higher = false;
            for (int w2 = w + 1; w2 < NWD; ++w2) higher |= MW(MK_VALID, w2, c) != 0u;
            unsigned gseed = 0, eseed = 0;
            if (!higher) {
                if (valid) {
                    const int i = 31 - __clz(valid);
                    if (static_cast<int>(s_angle[(lo + i) * S8 + c]) <= kp.s_start) gseed = 1u << i;  // :733
                } else if (w == 0) {
                    if (static_cast<int>(s_angle[c]) <= kp.s_start) eseed = 1u;  // accepted empty seed: labelled, not expanded (label.cpp:67-74)
                }
            }
            MW(MK_G, w, c) = gseed;
            MW(MK_KEEP, w, c) = eseed;
        }
        // the edge between the left neighbour's last column and my column 0: HA and HM words from the left halo
        for (int r = tid; r < 32 * NWD; r += NT) {  // whole warps: one per 32 rows
            const int w = r >> 5;
            const bool le = r < H && abs(static_cast<int>(s_haloL_angle[min(r, H - 1)]) - static_cast<int>(s_angle[min(r, H - 1) * S8])) <= kp.k_ground;
            const unsigned ha = __ballot_sync(0xffffffffu, le);
            if (lane == 0) {
                s_halo32[4 * 4 + w] = ha;
                s_halo32[3 * 4 + w] = ha & s_haloL_valid[w] & MW(MK_VALID, w, 0);
            }
        }
        }
        GS_PHASE(7);
        __syncthreads();  // the angle slab is free from here on (the link scan reuses it)
        GS_PHASE(8);
        // Ground label = reachability from the accepted seeds (the Labeler's BFS, label.cpp:48-104).  Nodes are the
        // vertical runs of every column; horizontal links unite runs of adjacent columns in a 16-bit union-find forest.
        // Unions go level by level (level k joins column groups of width 2^k across the boundaries b = 2^k - 1 mod 2^(k+1)),
        // which keeps the trees shallow; inside a level the links are dealt evenly to all threads through a prefix sum, so
        // every lane of a warp performs a union in every iteration.
        // Unions are organised so that most of them need no atomics.  Columns are cut into groups of GW = 2^GWL columns (at
        // least one group per warp).  Stage 1: a warp unites everything inside its group — no other warp touches those
        // trees.  Stage 2: the few links across group boundaries are united by all threads with compare-and-swap.  Links are
        // listed boundary by boundary: first the boundaries inside groups (ascending), then the group boundaries in level
        // order (low levels, small trees, first); q = (position of the boundary in that order) * NWD + word.
        int n_levels = 0;
        while ((1 << n_levels) < myCW) ++n_levels;
        int GWL = 4;
        while (GWL > 1 && ((myCW + (1 << GWL) - 1) >> GWL) < NT / 32) --GWL;
        const int GW = 1 << GWL;
        const int n_groups = (myCW + GW - 1) >> GWL;
        const int n_ext = (myCW - 1) >> GWL;      // group boundaries: c = GW - 1 mod GW, c <= myCW - 2
        const int n_int = (myCW - 1) - n_ext;     // boundaries inside groups
        const int n_q = (myCW - 1) * NWD;
        auto level_count = [=](int k) -> int {
            const int first = (1 << k) - 1;
            return (myCW - 2 >= first) ? ((myCW - 2 - first) >> (k + 1)) + 1 : 0;
        };
        auto boundary_of = [=](int p) -> int {  // p-th boundary in list order -> left column of the boundary
            if (p < n_int) {
                const int g = p / (GW - 1);
                return (g << GWL) + (p - g * (GW - 1));
            }
            p -= n_int;
            for (int k = GWL;; ++k) {
                const int nk = level_count(k);
                if (p < nk) return ((1 << k) - 1) + p * (2 << k);
                p -= nk;
            }
        };
        GS_REPEAT(7) {
            unsigned carry = 0;
            for (int q0 = 0; q0 < n_q; q0 += NT) {
                const int q = q0 + tid;
                unsigned todo = 0;
                if (q < n_q) {
                    const int pb = q / NWD, w = q - pb * NWD;
                    const int c = boundary_of(pb);
                    s_qcw[q] = static_cast<uint16_t>((c << 2) | w);
                    const unsigned hm = MW(MK_HM, w, c);
                    if (hm) {
                        // one union per pair of runs: skip link r when link r-1 joins the same two runs
                        const unsigned x = hm & MW(MK_V, w, c) & MW(MK_V, w, c + 1);
                        unsigned xprev = 0;
                        if (w > 0) xprev = (MW(MK_HM, w - 1, c) & MW(MK_V, w - 1, c) & MW(MK_V, w - 1, c + 1)) >> 31;
                        todo = hm & ~((x << 1) | xprev);
                    }
                    s_todo[q] = todo;
                }
                // block-wide exclusive prefix sum of the link counts
                const unsigned cnt = __popc(todo);
                unsigned incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) s_scan[warp] = incl;
                __syncthreads();
                if (warp == 0) {
                    unsigned ws = lane < NT / 32 ? s_scan[lane] : 0u;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned t = __shfl_up_sync(0xffffffffu, ws, d);
                        if (lane >= d) ws += t;
                    }
                    s_scan[32 + lane] = ws;  // inclusive over warps
                }
                __syncthreads();
                const unsigned before = carry + (warp ? s_scan[32 + warp - 1] : 0u) + incl - cnt;
                if (q < n_q) s_excl[q] = static_cast<uint16_t>(before);
                carry += s_scan[32 + NT / 32 - 1];
                __syncthreads();
            }
            if (tid == 0) s_excl[n_q] = static_cast<uint16_t>(carry);
            __syncthreads();
        }
        GS_PHASE(14);
        {
            auto at = [&](const void* ptr) -> unsigned { return smem_base32 + static_cast<unsigned>(reinterpret_cast<const unsigned char*>(ptr) - smem_raw); };
            const unsigned a_task = at(s_task), a_qcw = at(s_qcw), a_par = at(s_par);
            const unsigned a_runs = at(s_m32 + static_cast<size_t>(MK_S) * NWD * CW);  // MW(MK_S, w, c) = [w * CW + c]
            // link t of the list -> the forest nodes of the two runs it joins
            auto link_nodes = [&](int t_rel, int& na, int& nb) {
                const unsigned tk = lds_u16(a_task + 2 * t_rel);
                const int i = tk & 31;
                const int cw = lds_u16(a_qcw + 2 * (tk >> 5)), w = cw & 3, c = cw >> 2;
                const unsigned below = 0xffffffffu >> (31 - i);
                // run starts of (32w + i, c) and (32w + i, c + 1): highest run-start bit at or below the row
                int wa = w, wb = w;
                unsigned ma = lds_u32(a_runs + 4 * (w * CW + c)) & below, mb = lds_u32(a_runs + 4 * (w * CW + c + 1)) & below;
                while (ma == 0u) ma = lds_u32(a_runs + 4 * (--wa * CW + c));
                while (mb == 0u) mb = lds_u32(a_runs + 4 * (--wb * CW + c + 1));
                na = run_node(c, 32 * wa + 31 - __clz(ma));
                nb = run_node(c + 1, 32 * wb + 31 - __clz(mb));
            };
            auto write_tasks = [&](int qb, int qe, int base) {
                for (int q = qb + tid; q < qe; q += NT) {
                    unsigned word = s_todo[q];
                    uint16_t* dst = s_task + (static_cast<int>(s_excl[q]) - base);
                    while (word) {
                        *dst++ = static_cast<uint16_t>((q << 5) | (__ffs(word) - 1));
                        word &= word - 1;
                    }
                }
            };
            const int task_cap = static_cast<int>((smem_raw + off_masks - reinterpret_cast<unsigned char*>(s_task)) / 2);
            const int n_tasks = s_excl[n_q];
            if (n_tasks <= task_cap && !(A.stats & 4)) {
                // the whole list fits: warp-synchronous unions without atomics
                write_tasks(0, n_q, 0);
                __syncthreads();
                auto warp_range = [&](int ta, int tb) {  // the whole warp unites the links [ta, tb), 32 at a time
                    for (int t0 = ta; t0 < tb; t0 += 32) {
                        const bool act = t0 + lane < tb;
                        int na = 0, nb = 0;
                        if (act) link_nodes(t0 + lane, na, nb);
                        uf16s_unite_warp(a_par, act, na, nb);
                    }
                };
                for (int g = warp; g < n_groups; g += NT / 32) {
                    const int p0 = g * (GW - 1), p1 = min(p0 + GW - 1, n_int);
                    warp_range(s_excl[p0 * NWD], s_excl[p1 * NWD]);
                }
                __syncthreads();
                // stage 2: the links of the group boundaries (a few per cent of all), one per thread, with compare-and-swap
                for (int t = static_cast<int>(s_excl[n_int * NWD]) + tid; t < n_tasks; t += NT) {
                    int na, nb;
                    link_nodes(t, na, nb);
                    uf16s_unite_cas(a_par, na, nb);
                }
                __syncthreads();
            } else {
                // more links than the list holds (or GS_OPT_DEBUG_CAS_UNIONS): chunks of q, thread t takes the links t, t + NT, ...
                // of a chunk and unites with compare-and-swap
                int qb = 0;
                while (qb < n_q) {  // uniform
                    const int base = s_excl[qb];
                    int qe = n_q;
                    if (n_tasks - base > task_cap) {
                        int ql = qb + 1, qh = n_q;  // largest qe with excl[qe] - base <= cap (a word holds at most 32 <= cap links)
                        while (ql < qh) {
                            const int mid = (ql + qh + 1) >> 1;
                            if (static_cast<int>(s_excl[mid]) - base <= task_cap) ql = mid;
                            else qh = mid - 1;
                        }
                        qe = ql;
                    }
                    write_tasks(qb, qe, base);
                    __syncthreads();
                    const int n_t = static_cast<int>(s_excl[qe]) - base;
                    for (int t = tid; t < n_t; t += NT) {
                        int na, nb;
                        link_nodes(t, na, nb);
                        uf16s_unite_cas(a_par, na, nb);
                    }
                    __syncthreads();
                    qb = qe;
                }
            }
        }
        GS_PHASE(15);
        for (int u = tid; u < n_units; u += NT) {
            const int w = u / myCW, c = u - w * myCW;
            const unsigned seed = MW(MK_G, w, c);
            if (seed) {
                const int i = 31 - __clz(seed);
                int wa = w;
                unsigned ma = MW(MK_S, w, c) & (0xffffffffu >> (31 - i));
                while (ma == 0u) ma = MW(MK_S, --wa, c);
                uf16_flag_root(s_par, run_node(c, 32 * wa + 31 - __clz(ma)));
            }
        }
        __syncthreads();
        GS_PHASE(9);
        // Slab boundaries (and the column wrap of Label::WrapCols): the CTAs publish the ground masks of their two boundary
        // columns, import their neighbours' and flag the runs that ground reaches across the boundary, until a round in
        // which nobody flagged anything.  One cluster barrier per round (publish slots are double-buffered).
        uint16_t* s_broot = s_qcw + static_cast<size_t>(CW) * NWD + 2;  // [2][Hp] root of the run of (row, boundary column), 0xffff = empty pixel
        unsigned* s_pub = s_scan;                                        // [2 parities][2 columns][4 words]
        unsigned* s_chg = s_flags;                                       // [2 parities][MAX_CL], pushed by every CTA of the cluster
        for (int t = tid; t < 2 * H; t += NT) {
            const int b = t >= H ? 1 : 0, r = t - b * H;
            const int c = b ? myCW - 1 : 0;
            const int w = r >> 5, i = r & 31;
            unsigned root = 0xffffu;
            if ((MW(MK_VALID, w, c) >> i) & 1u) {
                int wa = w;
                unsigned ma = MW(MK_S, w, c) & (0xffffffffu >> (31 - i));
                while (ma == 0u) ma = MW(MK_S, --wa, c);
                root = static_cast<unsigned>(uf16_find(s_par, run_node(c, 32 * wa + 31 - __clz(ma))));
            }
            s_broot[b * Hp + r] = static_cast<uint16_t>(root);
        }
        __syncthreads();
        {
            int changed_prev = 1;  // round 0 always runs
            for (int round = 0;; ++round) {
                const int pk = round & 1;
                if (warp < 2) {  // publish: warp 0 = my column 0, warp 1 = my last column
                    for (int w = 0; w < NWD; ++w) {
                        const int r = 32 * w + lane;
                        bool gnd = false;
                        if (r < H) {
                            const unsigned root = s_broot[warp * Hp + r];
                            gnd = root != 0xffffu && (s_par[root] & 0x8000u) != 0;
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, gnd);
                        if (lane == 0) {
                            s_pub[(pk * 2 + warp) * 4 + w] = m;
                            if constexpr (SOFT) mail->pub[rank][pk][warp][w] = m;
                        }
                    }
                } else if (warp == 2 && lane < CL) {
                    if constexpr (SOFT) {
                        if (lane == 0) mail->chg[pk][rank] = static_cast<unsigned>(changed_prev);
                    } else {
                        cluster.map_shared_rank(s_chg, lane)[pk * MAX_CL + rank] = static_cast<unsigned>(changed_prev);
                    }
                }
                csync();
                unsigned anyone = 0;
                if constexpr (SOFT) {
                    for (int k = 0; k < CL; ++k) anyone |= __ldcg(&mail->chg[pk][k]);
                } else {
                    for (int k = 0; k < CL; ++k) anyone |= s_chg[pk * MAX_CL + k];
                }
                if (!anyone) break;
                bool changed = false;
                if (warp < 2) {
                    // warp 0: my column 0 against the left neighbour's last column (link = its HM, static halo slot 3)
                    // warp 1: my last column against the right neighbour's column 0 (link = my HM)
                    const int nb_rank = warp == 0 ? left_rank : right_rank;
                    const unsigned* their_pub = SOFT ? &mail->pub[nb_rank][pk][warp == 0 ? 1 : 0][0]
                                                     : cluster.map_shared_rank(s_pub, nb_rank) + (pk * 2 + (warp == 0 ? 1 : 0)) * 4;
                    const int c = warp == 0 ? 0 : myCW - 1;
                    for (int w = 0; w < NWD; ++w) {
                        const unsigned theirs = SOFT ? __ldcg(their_pub + w) : their_pub[w];
                        const unsigned mine = s_pub[(pk * 2 + warp) * 4 + w];
                        const unsigned link = warp == 0 ? s_halo32[3 * 4 + w] : MW(MK_HM, w, c);
                        if (lane == 0) s_halo32[(warp == 0 ? 2 : 1) * 4 + w] = theirs;  // kept for the label stage (gl / gr)
                        const unsigned in = theirs & link & ~mine;  // ground reaches these pixels of mine across the boundary
                        if ((in >> lane) & 1u) {
                            const unsigned root = s_broot[warp * Hp + 32 * w + lane];
                            unsigned* wp = reinterpret_cast<unsigned*>(s_par) + (root >> 1);
                            const int sh = (root & 1u) * 16;
                            const unsigned old = atomicOr(wp, 0x8000u << sh);
                            changed |= ((old >> sh) & 0x8000u) == 0;
                        }
                    }
                }
                changed_prev = __syncthreads_or(changed ? 1 : 0);
            }
        }
        GS_PHASE(10);
        // read the ground mask of every unit out of the forest, two runs at a time (lock-step walks)
        GS_REPEAT(4) {
            auto at = [&](const void* ptr) -> unsigned { return smem_base32 + static_cast<unsigned>(reinterpret_cast<const unsigned char*>(ptr) - smem_raw); };
            const unsigned a_par = at(s_par);
            const unsigned a_runs = at(s_m32 + static_cast<size_t>(MK_S) * NWD * CW), a_valid = at(s_m32 + static_cast<size_t>(MK_VALID) * NWD * CW);
            for (int u = tid; u < n_units; u += NT) {
                const int w = u / myCW, c = u - w * myCW;
                const unsigned valid = lds_u32(a_valid + 4 * (w * CW + c));
                unsigned runs = lds_u32(a_runs + 4 * (w * CW + c));
                unsigned g = 0;
                // pixels of a run that started in a lower word: the non-empty pixels below the first run start of this word
                unsigned head = valid & (runs ? ((runs & (0u - runs)) - 1u) : 0xffffffffu);
                auto next = [&](int& nd, unsigned& mk) -> bool {
                    if (head) {
                        int wa = w;
                        unsigned ma = 0;
                        while (ma == 0u) ma = lds_u32(a_runs + 4 * (--wa * CW + c));
                        nd = run_node(c, 32 * wa + 31 - __clz(ma));
                        mk = head;
                        head = 0;
                        return true;
                    }
                    if (runs) {
                        const unsigned bit = runs & (0u - runs);
                        runs ^= bit;
                        const unsigned upto_next = runs ? ((runs & (0u - runs)) - 1u) : 0xffffffffu;
                        nd = run_node(c, 32 * w + __ffs(bit) - 1);
                        mk = valid & upto_next & ~(bit - 1u);
                        return true;
                    }
                    return false;
                };
                int n0 = 0, n1 = 0;
                unsigned m0 = 0, m1 = 0;
                while (next(n0, m0)) {
                    if (!next(n1, m1)) {
                        n1 = n0;
                        m1 = 0;
                    }
                    unsigned e0, e1;
                    uf16s_root_entries2(a_par, n0, n1, e0, e1);
                    if (e0 & 0x8000u) g |= m0;
                    if (e1 & 0x8000u) g |= m1;
                    // nothing unites any more: hang both runs directly under their roots, which shortens the walks of the runs
                    // below them that other lanes still have to do (a reader sees the old parent or the root: both are ancestors)
                    if (READOUT_COMPRESS) {
                        const unsigned r0 = e0 & 0x7fffu, r1 = e1 & 0x7fffu;
                        if (r0 != static_cast<unsigned>(n0)) sts_u16(a_par + 2 * n0, r0);
                        if (r1 != static_cast<unsigned>(n1)) sts_u16(a_par + 2 * n1, r1);
                    }
                }
                MW(MK_G, w, c) = g;
            }
        }
        __syncthreads();

        GS_PHASE(11);
        // ================================================================================ D: label image, dilation, erase
        // L = G | accepted empty seeds | empty pixels 4-adjacent to a ground pixel u with |angle(u) - angle(p)| <= K
        const Mask<NW> Hmask = Mask<NW>::low(H);
        for (int c = tid; c < myCW; c += NT) {
            const Mask<NW> valid = ld_mask32<NW>(s_m32, MK_VALID, NWD, CW, c);
            const Mask<NW> g = ld_mask32<NW>(s_m32, MK_G, NWD, CW, c);
            const Mask<NW> va = ld_mask32<NW>(s_m32, MK_VA, NWD, CW, c);
            const Mask<NW> ha = ld_mask32<NW>(s_m32, MK_HA, NWD, CW, c);
            Mask<NW> gl, hal, gr;
            if (c == 0) {
                gl.w0 = s_halo32[2 * 4 + 0] | (NWD > 1 ? static_cast<unsigned long long>(s_halo32[2 * 4 + 1]) << 32 : 0ull);
                hal.w0 = s_halo32[4 * 4 + 0] | (NWD > 1 ? static_cast<unsigned long long>(s_halo32[4 * 4 + 1]) << 32 : 0ull);
                if constexpr (NW == 2) {
                    gl.w1 = (NWD > 2 ? s_halo32[2 * 4 + 2] : 0u) | (NWD > 3 ? static_cast<unsigned long long>(s_halo32[2 * 4 + 3]) << 32 : 0ull);
                    hal.w1 = (NWD > 2 ? s_halo32[4 * 4 + 2] : 0u) | (NWD > 3 ? static_cast<unsigned long long>(s_halo32[4 * 4 + 3]) << 32 : 0ull);
                }
            } else {
                gl = ld_mask32<NW>(s_m32, MK_G, NWD, CW, c - 1);
                hal = ld_mask32<NW>(s_m32, MK_HA, NWD, CW, c - 1);
            }
            if (c == myCW - 1) {
                gr.w0 = s_halo32[1 * 4 + 0] | (NWD > 1 ? static_cast<unsigned long long>(s_halo32[1 * 4 + 1]) << 32 : 0ull);
                if constexpr (NW == 2) gr.w1 = (NWD > 2 ? s_halo32[1 * 4 + 2] : 0u) | (NWD > 3 ? static_cast<unsigned long long>(s_halo32[1 * 4 + 3]) << 32 : 0ull);
            } else {
                gr = ld_mask32<NW>(s_m32, MK_G, NWD, CW, c + 1);
            }
            const Mask<NW> eseed = ld_mask32<NW>(s_m32, MK_KEEP, NWD, CW, c);
            const Mask<NW> near = (g.shr(1) & va) | (g & va).shl(1) | (gl & hal) | (gr & ha);
            const Mask<NW> L = (g | eseed | (~valid & near)) & Hmask;
            // dilated(r) = OR_k L(r + off_k)   (effective taps of cv::dilate with GetUniformKernel's bytes)
            Mask<NW> D = (kp.dil0 >= 0) ? L.shr(kp.dil0) : L.shl(-kp.dil0);
            if (kp.n_dil > 1) D = D | ((kp.dil1 >= 0) ? L.shr(kp.dil1) : L.shl(-kp.dil1));
            const Mask<NW> keep = valid & (kp.erase_ground ? ~D : D) & Hmask;  // alfa_gs.cpp:751-764
            st_mask32<NW>(s_m32, MK_KEEP, NWD, CW, c, keep);
        }
        __syncthreads();
        // transpose the KEEP words to row-major 32-column groups, count rows
        for (int job = warp; job < NG * NWD; job += NT / 32) {
            const int gidx = job / NWD, w = job - gidx * NWD;
            const int c = gidx * 32 + lane;
            const unsigned kw = c < myCW ? MW(MK_KEEP, w, c) : 0u;
            const int nrow = min(32, H - 32 * w);
            for (int i = 0; i < nrow; ++i) {
                const unsigned m = __ballot_sync(0xffffffffu, (kw >> i) & 1u);
                if (lane == 0) s_rowmask[(32 * w + i) * NG + gidx] = m;
            }
        }
        __syncthreads();
        for (int r = tid; r < H; r += NT) {
            unsigned run = 0;
            for (int gidx = 0; gidx < NG; ++gidx) {
                s_grp[r * NG + gidx] = run;
                run += __popc(s_rowmask[r * NG + gidx]);
            }
            if constexpr (SOFT) {
                s_rowcnt_all[rank * Hp + r] = run;
                mail->rowcnt[rank][r] = run;
            } else {
                for (int k = 0; k < CL; ++k) cluster.map_shared_rank(s_rowcnt_all, k)[rank * Hp + r] = run;
            }
        }
        // the frame this cluster takes next: rank 0 draws it from the queue all launches of the batch share and publishes it
        if (rank == 0 && tid == NT - 1) {
            const unsigned nf = static_cast<unsigned>(A.queue_base) + atomicAdd(A.queue, 1u);
            if constexpr (SOFT) mail->next_frame = nf;
            else
                for (int k = 0; k < CL; ++k) cluster.map_shared_rank(s_misc, k)[5] = nf;
        }
#undef MW
        GS_PHASE(12);
        csync();  // #: row counts of every slab (and the next frame) everywhere
        if constexpr (SOFT) {
            for (int t = tid; t < CL * H; t += NT) {
                const int k = t / H, r = t - k * H;
                if (k != rank) s_rowcnt_all[k * Hp + r] = __ldcg(&mail->rowcnt[k][r]);
            }
            if (tid == 0) s_misc[5] = __ldcg(&mail->next_frame);
            __syncthreads();
        }
        if (warp == 0) {
            // rowbase[r] = points in rows above (all slabs) + points of row r in the slabs left of mine
            unsigned carry = 0;
            for (int r0 = 0; r0 < H; r0 += 32) {
                const int r = r0 + lane;
                unsigned tot = 0, left = 0;
                if (r < H)
                    for (int k = 0; k < CL; ++k) {
                        const unsigned v = s_rowcnt_all[k * Hp + r];
                        tot += v;
                        if (k < rank) left += v;
                    }
                unsigned incl = tot;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (r < H) s_rowbase[r] = carry + incl - tot + left;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0 && rank == 0 && A.counts_out) A.counts_out[f] = carry;
        }
        __syncthreads();

        GS_PHASE(13);
        // the next frame of this cluster: pull the sampled points of its block dealing (P0) into L2 while the output streams out
        const int fn = static_cast<int>(min(s_misc[5], static_cast<unsigned>(A.frames)));
        f_next = fn;
        long long begn = 0;
        int Nn = 0;
        if (fn < A.frames) {  // uniform
            begn = A.offsets[fn];
            Nn = static_cast<int>(A.offsets[fn + 1] - begn);
            for (int i = 32 + 64 * tid; i < Nn; i += 64 * NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.points + begn + i));
        }
        // ================================================================================ B: no-ground image + cloud
        {
            uint16_t* res = A.res_out ? A.res_out + static_cast<size_t>(f) * kp.P : nullptr;
            float4* cloud = A.xyz_out ? A.xyz_out + static_cast<size_t>(f) * kp.P : nullptr;
            const bool range_as_double = kp.variant == GS_VARIANT_CLOUD2RANGE;
#if GS_EXP_DOUBLE == 3
            for (int rep_ = 0; rep_ < 2; ++rep_)
#endif
            for (int pair = warp; pair < H * NG; pair += NT / 32) {  // one warp per (row, 32-column group)
                const int r = SPEC ? pair / LC.NG : static_cast<int>(__umulhi(static_cast<unsigned>(pair), A.ng_magic));
                const int gidx = pair - r * NG;
                const int cc = gidx * 32 + lane;
                if ((SPEC && CWT % 32 == 0) || cc < myCW) {
                    const unsigned m = s_rowmask[pair];
                    const bool keep = (m >> lane) & 1u;
                    const unsigned enc = s_range[r * S16 + cc];
                    if (res) res[r * W + col0 + cc] = keep ? static_cast<uint16_t>(enc) : static_cast<uint16_t>(0);
                    if (keep && cloud) {
                        const unsigned slot = s_rowbase[r] + s_grp[pair] + __popc(m & ((1u << lane) - 1u));
                        const double rng_d = decode_range_d(kp, enc);
                        const double rng = range_as_double ? rng_d : static_cast<double>(__double2float_rn(rng_d));  // alfa_gs.cpp:1027,1045 / range2cloud_node.cpp:76
                        const double ca = s_cos_alt[r], sa = s_sin_alt[r];
                        const double caz = s_cos_az[cc], saz = s_sin_az[cc];
                        float4 o;
                        o.x = __double2float_rn(__dmul_rn(__dmul_rn(ca, caz), rng));  // :1055
                        o.y = __double2float_rn(__dmul_rn(__dmul_rn(ca, saz), rng));  // :1056
                        o.z = __double2float_rn(__dmul_rn(sa, rng));                  // :1057
                        o.w = 1.0f;
                        cloud[slot] = o;
                    }
                }
            }
        }
        // ... and build its block list once the output is on its way (measured: issuing the sample loads before the output
        // loop delays every warp's start and costs more than it hides)
        if (fn < A.frames) {  // uniform
            float2 deal_xy[2];
            deal_load(A.points + begn, Nn, 0, deal_xy);
            deal_vote(Nn, 0, deal_xy);
            deal_rest(A.points + begn, Nn);
        }
        beg_cur = begn;
        N_cur = Nn;
        GS_PHASE(0);
        __syncthreads();  // the slabs are cleared at the top of the next frame
    }

    if ((A.stats & 1)) {
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&A.cnt->dropped, static_cast<unsigned long long>(s_stat[ST_DROPPED]));
            atomicAdd(&A.cnt->near_edge_project, static_cast<unsigned long long>(s_stat[ST_NEAR]));
            atomicAdd(&A.cnt->near_edge_angle, static_cast<unsigned long long>(s_stat[ST_NEAR_ANGLE]));
            atomicAdd(&A.cnt->remote, static_cast<unsigned long long>(s_stat[ST_REMOTE]));
            atomicAdd(&A.cnt->slow_path, static_cast<unsigned long long>(s_stat[ST_FALLBACK]));
            atomicAdd(&A.cnt->fast_mismatch, static_cast<unsigned long long>(s_stat[ST_MISMATCH]));
        }
    }
    if constexpr (!SOFT) cluster.sync();  // no CTA exits while a neighbour may still address its shared memory
}

// ====================================================================================================================
// host side
// ====================================================================================================================
namespace {

bool layout_for(const KParams& kp, int CL, int NW, Layout* L) {
    const int W = kp.W, H = kp.H;
    const int CW = (W + CL - 1) / CL;
    if (CW < 2 && W >= 2) return false;
    if (W - (CL - 1) * CW <= 0) return false;  // the last CTA must own at least one column
    *L = layout_hcw(H, CW, NW);
    L->CL = CL;
    return L->ok;
}

struct FusedPlan {
    Layout lay;
    int NW;
    int max_clusters;
};

}  // namespace

namespace {
// the instantiation that serves (NW, threads, H, columns per CTA): specialised for the BASELINE sensor shapes, generic otherwise
const void* select_kernel(int NW, int nt, int H, int CW, int CL, int W, bool soft) {
#define GS_KF(...) (soft ? reinterpret_cast<const void*>(k_frame<__VA_ARGS__, true>) : reinterpret_cast<const void*>(k_frame<__VA_ARGS__, false>))
    if (nt == 1024 && W == CL * CW) {
        if (NW == 1 && H == 64 && CW == 512 && CL == 4) return GS_KF(1, 1024, 64, 512, 4);   // C2 / C4: 64 x 2048 on 4 CTAs
        if (NW == 1 && H == 64 && CW == 450 && CL == 4) return GS_KF(1, 1024, 64, 450, 4);   // the reference's Alfa_GS default: 64 x 1800 (alfa_gs.cpp:16-38)
        if (NW == 1 && H == 64 && CW == 500 && CL == 4) return GS_KF(1, 1024, 64, 500, 4);   // the reference's cloud2range default: 64 x 2000 (cloud2range_node.cpp:274-329)
        if (NW == 1 && H == 16 && CW == 900 && CL == 2) return GS_KF(1, 1024, 16, 900, 2);   // C1: 16 x 1800 on 2 CTAs
        if (NW == 2 && H == 128 && CW == 256 && CL == 8) return GS_KF(2, 1024, 128, 256, 8);  // C3 / C5: 128 x 2048 on 8 CTAs
    }
    if (nt == 1024) return NW == 1 ? GS_KF(1, 1024, 0, 0, 0) : GS_KF(2, 1024, 0, 0, 0);
    if (soft) return nullptr;  // the two-CTAs-per-SM experiment exists with thread-block clusters only
    return NW == 1 ? reinterpret_cast<const void*>(k_frame<1, 512, 0, 0, 0, false>) : reinterpret_cast<const void*>(k_frame<2, 512, 0, 0, 0, false>);
#undef GS_KF
}
}  // namespace

int fused_configure(FusedState& fs, const KParams& kp, const Luts& lut, const SgTaps& sg, bool smooth, bool repair, int device) {
    fs.kp = kp;
    fs.lut = lut;
    fs.sg = sg;
    fs.smooth = smooth;
    fs.repair = repair;
    fs.configured = false;
    // what the fused kernel covers; everything else runs on the staged kernels
    if (kp.pixel_policy != GS_PIXEL_LAST && kp.pixel_policy != GS_PIXEL_NEAREST) return 0;
    if (kp.H > MAX_H || kp.H < 2 || kp.W < 2 || (kp.H & (kp.H - 1)) != 0) return 0;  // beam counts that are powers of two (16 / 32 / 64 / 128)
    const int NW = kp.H <= 64 ? 1 : 2;
    int max_smem = 0, n_sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return static_cast<int>(e);
    // Pick the cluster size (1..8 CTAs per frame, not only powers of two) that keeps the most SMs busy: clusters are placed
    // inside one GPC, so e.g. clusters of 4 strand 16 of a B200's 148 SMs while clusters of 6 strand 4.  Ties go to the
    // smaller cluster (fewer cross-CTA exchanges).  GS_FUSED_CLUSTER=n forces a size, GS_FUSED_THREADS=512 the two-CTAs-per-SM shape.
    int max_smem_sm = 0;
    e = cudaDeviceGetAttribute(&max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    if (e != cudaSuccess) return static_cast<int>(e);
    const char* mode = getenv("GS_FUSED_THREADS");
    const char* force_cl = getenv("GS_FUSED_CLUSTER");
    const int nt = (mode && atoi(mode) == 512) ? 512 : 1024;
    const size_t smem_cap = nt == 512 ? static_cast<size_t>(max_smem_sm) / 2 - 1024 : static_cast<size_t>(max_smem);
    Layout lay;
    int max_clusters = 0, best_sms = 0;
    for (int CL = 1; CL <= MAX_CL; ++CL) {
        if (force_cl && atoi(force_cl) != CL) continue;
        Layout l;
        if (!layout_for(kp, CL, NW, &l) || l.total > smem_cap) continue;
        const void* kern = select_kernel(NW, nt, kp.H, l.CW, CL, kp.W, false);
        // the attribute is per function, not per context: always ask for the device maximum
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>((nt == 512 ? 2 : 1) * n_sms / CL * CL));
        cfg.blockDim = dim3(nt);
        cfg.dynamicSmemBytes = l.total;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = static_cast<unsigned>(CL);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        if (getenv("GS_DEBUG_PLAN")) fprintf(stderr, "gs fused candidate: cluster %d, %d columns per CTA, %zu B smem, %d co-resident clusters = %d CTAs\n", CL, l.CW, l.total, n, n * CL);
        if (n * CL > best_sms) {
            best_sms = n * CL;
            max_clusters = n;
            lay = l;
            fs.kernel = kern;
        }
    }
    if (max_clusters < 1) return 0;  // no cluster shape fits here: staged kernels
    // Thread-block clusters are placed inside one GPC, which strands SMs (clusters of 4: 132 of 148).  SOFTWARE clusters — the same
    // kernel with its cluster exchanges through global memory, launched cooperatively beside the cluster kernel — take the SMs that
    // are left (GS_FUSED_MODE=hybrid, the default), or all of them (=soft); =hw keeps thread-block clusters only.
    const char* mode_env = getenv("GS_FUSED_MODE");
    int fmode = FUSED_MODE_HYBRID;
    if (mode_env && !strcmp(mode_env, "hw")) fmode = FUSED_MODE_HW;
    if (mode_env && !strcmp(mode_env, "soft")) fmode = FUSED_MODE_SOFT;
    fs.kernel_soft = nt == 1024 ? select_kernel(NW, nt, kp.H, lay.CW, lay.CL, kp.W, true) : nullptr;
    fs.soft_clusters = 0;
    if (fs.kernel_soft && fmode != FUSED_MODE_HW) {
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (coop && cudaFuncSetAttribute(fs.kernel_soft, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fs.kernel_soft, nt, lay.total) == cudaSuccess && per_sm >= 1) {
            const int free_sms = fmode == FUSED_MODE_SOFT ? n_sms : n_sms - max_clusters * lay.CL;
            fs.soft_clusters = free_sms / lay.CL;
        } else {
            cudaGetLastError();
        }
    }
    if (fs.soft_clusters < 1) fmode = FUSED_MODE_HW;
    fs.mode = fmode;
    if (getenv("GS_DEBUG_PLAN")) fprintf(stderr, "gs fused plan: %d x %d image, %d threads, cluster %d, %d columns per CTA, %zu B smem, %d co-resident clusters on %d SMs, mode %d, %d software clusters\n",
                                         kp.H, kp.W, nt, lay.CL, lay.CW, lay.total, max_clusters, n_sms, fmode, fs.soft_clusters);
    fs.cluster = lay.CL;
    fs.threads = nt;
    fs.cols_per_cta = lay.CW;
    fs.smem_bytes = lay.total;
    fs.n_sms = n_sms;
    fs.max_clusters = max_clusters;
    // A lone frame (gs_process_frame: the reference's one-callback-per-cloud model) has nothing to share the GPU with: it is spread
    // over the widest cluster that fits (8 CTAs x 256 columns for 64 x 2048) instead of the batch plan's cluster.
    fs.lone_cluster = 0;
    fs.lone_kernel = nullptr;
    if (nt == 1024 && !force_cl && !getenv("GS_NO_LONE_PLAN")) {
        for (int CL = MAX_CL; CL > lay.CL; --CL) {
            Layout l;
            if (!layout_for(kp, CL, NW, &l) || l.total > smem_cap || l.CW < 64) continue;
            const void* kern = select_kernel(NW, nt, kp.H, l.CW, CL, kp.W, false);
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) != cudaSuccess) {
                cudaGetLastError();
                continue;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(static_cast<unsigned>(CL));
            cfg.blockDim = dim3(nt);
            cfg.dynamicSmemBytes = l.total;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = static_cast<unsigned>(CL);
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
                cudaGetLastError();
                continue;
            }
            fs.lone_cluster = CL;
            fs.lone_kernel = kern;
            break;
        }
    }
    if (!fs.d_const) {
        e = cudaMalloc(&fs.d_const, sizeof(KParams) + sizeof(GateConsts));
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    e = cudaMemcpy(fs.d_const, &kp, sizeof(KParams), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return static_cast<int>(e);
    {
        static_assert(sizeof(KParams) % 8 == 0, "GateConsts follows KParams in one allocation");
        GateConsts g;
        g.range_lo_in = kp.min_range + 1e-9;
        g.range_hi_in = kp.max_range - 1e-9;
        g.range_lo_out = kp.min_range - 1e-9;
        g.range_hi_out = kp.max_range + 1e-9;
        g.alt_lo = static_cast<float>(kp.min_angle);
        g.alt_hi = static_cast<float>(kp.max_angle);
        e = cudaMemcpy(static_cast<unsigned char*>(fs.d_const) + sizeof(KParams), &g, sizeof(g), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) return static_cast<int>(e);
    fs.configured = true;
    return 0;
}

bool fused_supported(const FusedState& fs) { return fs.configured && !fs.disabled; }

void fused_release(FusedState& fs) {
    cudaFree(fs.d_const);
    cudaFree(fs.d_rowbase);
    cudaFree(fs.d_mail);
    if (fs.stream2) cudaStreamDestroy(fs.stream2);
    if (fs.ev_fork) cudaEventDestroy(fs.ev_fork);
    if (fs.ev_join) cudaEventDestroy(fs.ev_join);
    fs.d_const = nullptr;
    fs.d_rowbase = nullptr;
    fs.d_mail = nullptr;
    fs.stream2 = nullptr;
    fs.ev_fork = fs.ev_join = nullptr;
    fs.rowbase_cap = 0;
    fs.configured = false;
}

long long fused_max_points() { return static_cast<long long>(MAX_BLOCKS) * BLK; }

int fused_run(FusedState& fs, const Launch& L, const float4* points, const long long* d_offsets, int frames, long long max_points,
              uint16_t* range_out, uint16_t* angle_out, uint16_t* res_out, float4* xyz_out, unsigned* counts_out, Counters* cnt,
              int stats) {
    if (frames <= 0) return 0;
    if (max_points > fused_max_points()) return static_cast<int>(cudaErrorInvalidValue);
    const KParams& kp = fs.kp;
    const int NW = kp.H <= 64 ? 1 : 2;
    const bool lone = frames == 1 && fs.lone_cluster > 0 && fs.lone_kernel && !fs.no_lone;  // one frame: the widest cluster that fits
    Layout lay;
    if (!layout_for(kp, lone ? fs.lone_cluster : fs.cluster, NW, &lay)) return static_cast<int>(cudaErrorInvalidConfiguration);
    // how the frames are shared out: clusters of the cluster kernel first, software clusters for what is left
    const int n_hw = lone ? 1 : (fs.mode == FUSED_MODE_SOFT ? 0 : std::min(frames, fs.max_clusters));
    const int n_soft = (lone || fs.mode == FUSED_MODE_HW) ? 0 : std::min(frames - n_hw, fs.soft_clusters);
    const int slots_all = fs.max_clusters + fs.soft_clusters;  // mailboxes / queue: sized once for the batch plan
    const int slots = lone ? 1 : slots_all;
    // remote-record scratch: every point of a frame may be remote in the worst case (shuffled clouds)
    const unsigned rcap = static_cast<unsigned>(std::max<long long>((max_points + 1023) / 1024 * 1024, 1024));
    const size_t need = static_cast<size_t>(slots) * lay.CL * (static_cast<size_t>(rcap) * sizeof(uint2) + sizeof(unsigned));
    if (fs.rowbase_cap < need) {
        cudaFree(fs.d_rowbase);
        fs.d_rowbase = nullptr;
        fs.rowbase_cap = 0;
        const cudaError_t e = cudaMalloc(&fs.d_rowbase, need);
        if (e != cudaSuccess) return static_cast<int>(e);
        fs.rowbase_cap = need;
    }
    // frame queue + the mailboxes of the software clusters: one allocation, zeroed before every batch
    const size_t mail_bytes = 128 + static_cast<size_t>(std::max(slots_all, 1)) * sizeof(SoftMail);
    if (!fs.d_mail) {
        cudaError_t e = cudaMalloc(&fs.d_mail, mail_bytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        if (fs.soft_clusters > 0) {
            // lowest priority: the cluster kernel is placed first (its clusters need whole TPC groups of a GPC), the software clusters
            // then take the SMs that are left
            int prio_lo = 0, prio_hi = 0;
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
            if ((e = cudaStreamCreateWithPriority(&fs.stream2, cudaStreamNonBlocking, prio_lo)) != cudaSuccess) return static_cast<int>(e);
            if ((e = cudaEventCreateWithFlags(&fs.ev_fork, cudaEventDisableTiming)) != cudaSuccess) return static_cast<int>(e);
            if ((e = cudaEventCreateWithFlags(&fs.ev_join, cudaEventDisableTiming)) != cudaSuccess) return static_cast<int>(e);
        }
    }
    {
        const cudaError_t e = cudaMemsetAsync(fs.d_mail, 0, lone ? 128 : mail_bytes, L.stream);  // a lone frame only needs the queue word
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    FusedArgs a;
    a.kp = kp;
    a.lut = fs.lut;
    a.CL = lay.CL;
    a.CW = lay.CW;
    a.S16 = lay.S16;
    a.S8 = lay.S8;
    a.NG = lay.NG;
    a.off_range = lay.off_range;
    a.off_tag = lay.off_tag;
    a.off_angle = lay.off_angle;
    a.off_masks = lay.off_masks;
    a.off_lut = lay.off_lut;
    a.off_misc = lay.off_misc;
    a.points = points;
    a.offsets = d_offsets;
    a.frames = frames;
    a.range_out = range_out;
    a.angle_out = angle_out;
    a.res_out = res_out;
    a.xyz_out = xyz_out;
    a.counts_out = counts_out;
    a.cnt = cnt;
    a.stats = stats;  // bit 0: verification + counters, bit 1: phase timing
    a.rcap = rcap;
    // fast-path constants (see project_lean2)
    {
        const double two_fb = static_cast<double>(1 << FB);
        const bool c2r = kp.variant == GS_VARIANT_CLOUD2RANGE;
        const double alt_scale = c2r ? static_cast<double>(kp.H) / kp.ang_span : 1.0 / kp.d_alt;  // cloud2range_node.cpp:436 / alfa_gs.cpp:275
        const double row0 = c2r ? kp.max_angle * alt_scale : static_cast<double>(kp.H - 1) + (kp.half_d_alt + kp.min_angle) / kp.d_alt;
        const double az_scale = static_cast<double>(kp.W) / 6.283185307179586476925286766559;   // = 1 / d_azimuth
        a.enc_scale = 65534.0 / kp.span;
        a.alt_to_fix = static_cast<float>(alt_scale * two_fb);
        a.az_to_fix = static_cast<float>(az_scale * two_fb);
        a.alt_lo = static_cast<float>(kp.min_angle);
        a.alt_hi = static_cast<float>(kp.max_angle);
        a.row0_fix = static_cast<int>(llround(row0 * two_fb));
        a.quarter_fix = kp.W << (FB - 2);
        a.col_half_fix = c2r ? 0 : (1 << (FB - 1));
        a.row_margin = static_cast<int>(std::ceil(1.0e-6 * alt_scale * two_fb)) + 8;   // 1e-6 rad >= 4x the altitude error budget
        a.col_margin = static_cast<int>(std::ceil(7.5e-7 * az_scale * two_fb)) + 8;    // 7.5e-7 rad >= 4x the azimuth error budget
        a.u_scale = a.enc_scale * static_cast<double>(1 << UB);
        a.u_c0 = (1.0 - kp.min_range * a.enc_scale) * static_cast<double>(1 << UB);
        {
            // altitude gates (alfa_gs.cpp:269) on rn(altitude * alt_to_fix): inside by more than 1e-6 rad (+ 2 units of rounding)
            const double glo = std::ceil((kp.min_angle + 1.0e-6) * alt_scale * two_fb) + 2.0;
            const double ghi = std::floor((kp.max_angle - 1.0e-6) * alt_scale * two_fb) - 2.0;
            a.alt_gate_lo = static_cast<int>(glo);
            a.alt_gate_span = ghi >= glo ? static_cast<int>(ghi - glo) : -1;
        }
        a.cw_magic = static_cast<unsigned>((0x100000000ull + lay.CW - 1) / lay.CW);
        a.ng_magic = static_cast<unsigned>((0x100000000ull + lay.NG - 1) / lay.NG);
        a.kp_dev = static_cast<const KParams*>(fs.d_const);
        a.gates_dev = reinterpret_cast<const GateConsts*>(static_cast<const unsigned char*>(fs.d_const) + sizeof(KParams));
        a.col_limit = c2r ? kp.W - 1 : kp.W;
        a.sg_ws = fs.smooth ? fs.sg.ws : 0;
        a.repair = fs.repair ? 1 : 0;
        for (int k = 0; k < 11; ++k) a.sg_t[k] = fs.sg.t[k];
        a.range_lo_in = kp.min_range + 1e-9;
        a.range_hi_in = kp.max_range - 1e-9;
        a.range_lo_out = kp.min_range - 1e-9;
        a.range_hi_out = kp.max_range + 1e-9;
        a.col_wrap = (!c2r && kp.col_policy == GS_COL_WRAP) ? 1 : 0;
        // the fixed-point coordinates need 4 * quarter_fix < 2^31 and |altitude cells| * 2^FB < 2^31
        a.use_fast = (kp.W <= 8191 && alt_scale * 1.6 * two_fb < 2.0e9 && std::fabs(row0) * two_fb < 2.0e9 &&
                      a.row_margin < (1 << (FB - 2)) && a.col_margin < (1 << (FB - 2))) ? 1 : 0;
    }
    a.rlist = reinterpret_cast<uint2*>(fs.d_rowbase);
    a.rcount = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(fs.d_rowbase) +
                                           static_cast<size_t>(slots) * lay.CL * rcap * sizeof(uint2));
    a.queue = reinterpret_cast<unsigned*>(fs.d_mail);
    a.mail = reinterpret_cast<SoftMail*>(reinterpret_cast<unsigned char*>(fs.d_mail) + 128);
    a.queue_base = n_hw + n_soft;
    const bool two = n_hw > 0 && n_soft > 0;
    if (two) {
        // the software clusters run beside the cluster kernel on a second stream, ordered after the work already queued on L.stream
        // (inputs, the memset above) and joined back into it below
        cudaError_t e;
        if ((e = cudaEventRecord(fs.ev_fork, L.stream)) != cudaSuccess) return static_cast<int>(e);
        if ((e = cudaStreamWaitEvent(fs.stream2, fs.ev_fork, 0)) != cudaSuccess) return static_cast<int>(e);
    }
    if (n_hw > 0) {
        a.first_frame = 0;
        a.slot_base = 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(n_hw * lay.CL));
        cfg.blockDim = dim3(fs.threads);
        cfg.dynamicSmemBytes = lay.total;
        cfg.stream = L.stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = static_cast<unsigned>(lay.CL);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        void* kargs[] = {&a};
        const cudaError_t e = cudaLaunchKernelExC(&cfg, lone ? fs.lone_kernel : fs.kernel, kargs);
        if (e != cudaSuccess) return static_cast<int>(e);
        ++*L.n_launches;
    }
    if (n_soft > 0) {
        cudaStream_t st = two ? fs.stream2 : L.stream;
        cudaError_t e;
        a.first_frame = n_hw;
        a.slot_base = n_hw > 0 ? fs.max_clusters : 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(n_soft * lay.CL));
        cfg.blockDim = dim3(fs.threads);
        cfg.dynamicSmemBytes = lay.total;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the software clusters wait for each other
        attr[0].val.cooperative = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        void* kargs[] = {&a};
        e = cudaLaunchKernelExC(&cfg, fs.kernel_soft, kargs);
        if (e != cudaSuccess) return static_cast<int>(e);
        ++*L.n_launches;
        if (two) {
            if ((e = cudaEventRecord(fs.ev_join, st)) != cudaSuccess) return static_cast<int>(e);
            if ((e = cudaStreamWaitEvent(L.stream, fs.ev_join, 0)) != cudaSuccess) return static_cast<int>(e);
        }
    }
    return 0;
}

}  // namespace gs
