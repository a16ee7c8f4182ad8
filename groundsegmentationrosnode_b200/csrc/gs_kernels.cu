// gs_kernels.cu — the hot path as hand-written CUDA for sm_100a.
//
// Stage map (reference file:line -> kernel):
//   K1  k_project + k_finalize     Alfa_GS::CreateRangeImage   src/alfa_gs.cpp:247-298 (+ lib/utils.h:8-43)
//                                  Cloud2RangeNode::CloudCb    src/cloud2range_node.cpp:396-455
//   K2  k_angle                    Alfa_GS::CreateAngleImg     src/alfa_gs.cpp:335-386
//       k_sg                       SavitskyGolaySmoothing      src/alfa_gs.cpp:641-713 (cv::filter2D)
//       k_repair                   RepairGaps (typed)          src/alfa_gs.cpp:300-333
//   K3  k_ccl_init/k_ccl_union     Label::LabelOneComponent    Labeler/label.cpp:48-104 as union-find
//       k_seeds, k_label           EraseBFS seed loop          src/alfa_gs.cpp:720-739
//   K4  k_erase                    GetUniformKernel+cv::dilate+erase  src/alfa_gs.cpp:747-764
//       k_backproject              Alfa_GS::CameraCb           src/alfa_gs.cpp:997-1074 / src/range2cloud_node.cpp:64-91
//
// Everything here is HBM/L2-bound integer and fp work: no tensor cores (nothing is a contraction).
// All kernels take a wave of independent frames (blockIdx.y = frame) so that one launch fills the 148 SMs.
// Parity rules: compile with -fmad=false; parity-critical arithmetic additionally uses explicit
// round-to-nearest intrinsics so it does not depend on the flag.
#include <math_constants.h>

#include <algorithm>

#include "gs_internal.h"
#include "gs_math.cuh"

namespace gs {

namespace {

constexpr int FLAG_GROUND = 1 << 30;
constexpr int IDX_MASK = FLAG_GROUND - 1;

// streaming 16-byte load: points are read exactly once, keep them out of L1 and first in line for L2 eviction
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) { return __ldcs(p); }

__device__ __forceinline__ double decode_range(const KParams& kp, unsigned enc) {
    // static_cast<double>(enc - 1) / 65534 * (max_range - min_range) + min_range   (alfa_gs.cpp:359-360, :1042-1045)
    const double norm = __ddiv_rn(static_cast<double>(static_cast<int>(enc) - 1), 65534.0);
    return __dadd_rn(__dmul_rn(norm, kp.span), kp.min_range);
}

// ------------------------------------------------------------------------------------------------
// K1 — spherical projection.  One thread per point; the pixel owner is decided by a packed 64-bit
// atomicMax: (point index << 32 | range) keeps the LAST point in input order like the reference's
// sequential overwrite (alfa_gs.cpp:291); the `nearest` policy stores the complement of
// (range << 32 | index) so that the same atomicMax keeps the smallest range, lowest index.
// ------------------------------------------------------------------------------------------------
struct Projected {
    int flat;      // row * W + col, or -1 when the point is dropped
    unsigned enc;  // 1..65535
    bool near_edge;
};

__device__ __forceinline__ Projected project_point(const KParams& kp, float fx, float fy, float fz) {
    Projected out;
    out.flat = -1;
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
    int flat;
    if (kp.variant == GS_VARIANT_CLOUD2RANGE) {
        if (col >= kp.W) return out;
        flat = row * kp.W + col;
    } else if (kp.col_policy == GS_COL_WRAP) {
        col %= kp.W;
        flat = row * kp.W + col;
    } else {
        flat = row * kp.W + col;  // Q2: col may equal W and alias (row+1, 0)
        if (flat >= kp.P) return out;
    }
    const double range_norm = __ddiv_rn(__dsub_rn(range, kp.min_range), kp.span);  // alfa_gs.cpp:288
    out.enc = static_cast<unsigned>(__double2int_rz(__dadd_rn(__dmul_rn(range_norm, 65534.0), 1.0)));  // :291
    out.flat = flat;
    return out;
}

template <bool STRIDE16>
__global__ void __launch_bounds__(256) k_project(const char* __restrict__ points, size_t stride,
                                                 const long long* __restrict__ offsets, KParams kp,
                                                 unsigned long long* __restrict__ keys, Counters* cnt, int stats) {
    const int f = blockIdx.y;
    const long long beg = offsets[f], end = offsets[f + 1];
    const long long i = beg + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= end) return;
    float fx, fy, fz;
    if (STRIDE16) {
        const float4 q = ld_stream_f4(reinterpret_cast<const float4*>(points) + i);
        fx = q.x; fy = q.y; fz = q.z;
    } else {
        const float* q = reinterpret_cast<const float*>(points + static_cast<size_t>(i) * stride);
        fx = q[0]; fy = q[1]; fz = q[2];
    }
    const Projected pr = project_point(kp, fx, fy, fz);
    if (stats) {
        if (pr.flat < 0) atomicAdd(&cnt->dropped, 1ull);
        else if (pr.near_edge) atomicAdd(&cnt->near_edge_project, 1ull);
    }
    if (pr.flat < 0) return;
    const unsigned long long idx = static_cast<unsigned long long>(i - beg);
    unsigned long long key;
    if (kp.pixel_policy == GS_PIXEL_NEAREST) key = ~((static_cast<unsigned long long>(pr.enc) << 32) | idx);
    else key = (idx << 32) | pr.enc;
    atomicMax(keys + static_cast<size_t>(f) * kp.P + pr.flat, key);
}

// keys -> u16 range image; keys are re-zeroed for the next wave.
__global__ void __launch_bounds__(256) k_finalize(unsigned long long* __restrict__ keys, uint16_t* __restrict__ range,
                                                  KParams kp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kp.P) return;
    const size_t i = static_cast<size_t>(blockIdx.y) * kp.P + p;
    const unsigned long long key = keys[i];
    unsigned enc = 0;
    if (key) {
        keys[i] = 0ull;
        enc = (kp.pixel_policy == GS_PIXEL_NEAREST) ? static_cast<unsigned>((~key) >> 32) : static_cast<unsigned>(key & 0xffffu);
    }
    range[i] = static_cast<uint16_t>(enc);
}

// ------------------------------------------------------------------------------------------------
// K2 — angle image.  The reference walks every column bottom->top carrying the previous non-empty
// pixel (alfa_gs.cpp:347-382).  Here every pixel finds that predecessor itself by scanning down its
// column (1-2 L1-resident loads in practice), which makes all pixels independent and all loads
// row-coalesced.  sinf/cosf of the row angle come from host LUTs; atan2f is the glibc clone.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_angle(const uint16_t* __restrict__ range, uint16_t* __restrict__ angle, KParams kp,
                                               Luts lut, Counters* cnt, int stats) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kp.P) return;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    const uint16_t* R = range + base;
    const unsigned enc = R[p];
    unsigned out = 0;
    if (enc) {
        const int r = p / kp.W, c = p - r * kp.W;
        int rr = r + 1;
        unsigned penc = 0;
        for (; rr < kp.H; ++rr) {
            penc = R[rr * kp.W + c];
            if (penc) break;
        }
        bool has_prev = penc != 0;
        float prev_r = 0.0f, sp = 0.0f, cp = 1.0f;  // initial state: prev_r = 0, prev_angle_y = 0 (alfa_gs.cpp:342)
        if (has_prev) {
            prev_r = __double2float_rn(decode_range(kp, penc));  // prev_r = range (float) :379
            sp = lut.sin_y[rr];
            cp = lut.cos_y[rr];
        } else if (c == 1) {
            // column-1 quirk (SURVEY.md A.2 iii): prev_col starts at 1 (:340), so the first pixel ever
            // processed, if it lies in column 1, is compared against the initial state.
            bool col0_empty = true;
            for (int q = 0; q < kp.H; ++q)
                if (R[q * kp.W]) { col0_empty = false; break; }
            has_prev = col0_empty;
        }
        if (has_prev) {
            const double rng = decode_range(kp, enc);
            // :366-367  |float*float - double*float| in double, narrowed to float
            const double dz = fabs(__dsub_rn(static_cast<double>(__fmul_rn(prev_r, sp)), __dmul_rn(rng, static_cast<double>(lut.sin_y[r]))));
            const double dx = fabs(__dsub_rn(static_cast<double>(__fmul_rn(prev_r, cp)), __dmul_rn(rng, static_cast<double>(lut.cos_y[r]))));
            const float alfa = gs_atan2f(__double2float_rn(dz), __double2float_rn(dx));  // :368
            const float scaled = __fmul_rn(alfa, 100.0f);                                 // :370
            out = static_cast<unsigned>(__float2int_rz(scaled));
            if (stats) {
                const float fr = scaled - floorf(scaled);
                if (fminf(fr, 1.0f - fr) < 1e-3f) atomicAdd(&cnt->near_edge_angle, 1ull);
            }
        }
    }
    angle[base + p] = static_cast<uint16_t>(out);
}

// Savitzky-Golay column filter = cv::filter2D(u16, ws x 1 float kernel, BORDER_REFLECT101): float32 accumulation in
// tap order, round-half-even, saturate (pinned against OpenCV in tests/golden).
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(256) k_sg(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, KParams kp, SgTaps taps) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kp.P) return;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    const int r = p / kp.W, c = p - r * kp.W, half = taps.ws / 2;
    float acc = 0.0f;
    for (int k = 0; k < taps.ws; ++k) {
        const int rr = reflect101(r + k - half, kp.H);
        acc = __fadd_rn(acc, __fmul_rn(taps.t[k], static_cast<float>(in[base + rr * kp.W + c])));
    }
    int v = __float2int_rn(acc);
    v = max(0, min(65535, v));
    out[base + p] = static_cast<uint16_t>(v);
}

// Typed RepairGaps: in-place, top-to-bottom dependency inside a column -> one thread per column.
__global__ void __launch_bounds__(128) k_repair(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, KParams kp, int step,
                                                float thr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kp.W) return;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    for (int r = 0; r < kp.H; ++r) out[base + r * kp.W + c] = in[base + r * kp.W + c];
    for (int r = 0; r < kp.H; ++r) {
        if (out[base + r * kp.W + c] != 0) continue;
        int counter = 0;
        float sum = 0.0f;
        for (int i = 1; i < step; ++i) {
            if (r - i < 0) continue;
            for (int j = 1; j < step; ++j) {
                if (r + j > kp.H - 1) continue;
                const float prev = out[base + (r - i) * kp.W + c];
                const float next = out[base + (r + j) * kp.W + c];
                if (prev > 0.001f && next > 0.001f && fabsf(__fsub_rn(prev, next)) < thr) {
                    sum = __fadd_rn(sum, __fadd_rn(prev, next));
                    counter += 2;
                }
            }
        }
        if (counter > 0) out[base + r * kp.W + c] = static_cast<uint16_t>(__float2int_rz(__fdiv_rn(sum, static_cast<float>(counter))));
    }
}

// ------------------------------------------------------------------------------------------------
// K3 — connected components by union-find (root = smallest flat index), replacing the Labeler's BFS.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* par, int i) {
    int p = par[i] & IDX_MASK;
    while (p != i) {
        i = p;
        p = par[i] & IDX_MASK;
    }
    return i;
}

__device__ __forceinline__ void uf_unite(int* par, int a, int b) {
    for (;;) {
        a = uf_find(par, a);
        b = uf_find(par, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(par + a, b);  // hang the larger root under the smaller one
        if (old == a) return;
        a = old;  // somebody re-parented a first: unite its new parent with b as well
    }
}

__global__ void __launch_bounds__(256) k_ccl_init(int* __restrict__ parent, KParams kp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kp.P) return;
    parent[static_cast<size_t>(blockIdx.y) * kp.P + p] = p;
}

__device__ __forceinline__ bool px_open(const uint16_t* R, const uint8_t* blocked, int p) {
    return R[p] != 0 && (blocked == nullptr || blocked[p] == 0);
}

// One thread per pixel links it with its upper and left neighbour (left of column 0 is column W-1:
// Label::WrapCols, label.cpp:33-45) when both are non-empty and |angle difference| <= k_edge
// (label.cpp:90-101 with the float threshold folded into an integer).
__global__ void __launch_bounds__(256) k_ccl_union(const uint16_t* __restrict__ range, const uint16_t* __restrict__ angle,
                                                   const uint8_t* __restrict__ blocked, int k_edge, int* __restrict__ parent,
                                                   KParams kp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kp.P) return;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    const uint16_t* R = range + base;
    const uint16_t* A = angle + base;
    const uint8_t* B = blocked ? blocked + base : nullptr;
    int* par = parent + base;
    if (!px_open(R, B, p)) return;
    const int r = p / kp.W, c = p - r * kp.W;
    const int a = A[p];
    if (r > 0) {
        const int q = p - kp.W;
        if (px_open(R, B, q) && abs(a - static_cast<int>(A[q])) <= k_edge) uf_unite(par, p, q);
    }
    if (kp.W > 1) {
        const int q = (c > 0) ? p - 1 : p + kp.W - 1;
        if (px_open(R, B, q) && abs(a - static_cast<int>(A[q])) <= k_edge) uf_unite(par, p, q);
    }
}

// Seed loop of EraseBFS (alfa_gs.cpp:720-739): per column the bottom-most non-empty pixel (stopping at row 0),
// accepted when its angle <= s_start; accepted non-empty seeds flag their component's root as ground.
__global__ void __launch_bounds__(128) k_seeds(const uint16_t* __restrict__ range, const uint16_t* __restrict__ angle,
                                               int* __restrict__ parent, int* __restrict__ seed_row, uint8_t* __restrict__ eseed,
                                               KParams kp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kp.W) return;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    const uint16_t* R = range + base;
    int r = kp.H - 1;
    while (r > 0 && R[r * kp.W + c] == 0) --r;  // :722-725
    const int p = r * kp.W + c;
    const bool accepted = static_cast<int>(angle[base + p]) <= kp.s_start;  // :733
    const bool valid = R[p] != 0;
    if (accepted && valid) {
        int* par = parent + base;
        atomicOr(par + uf_find(par, p), FLAG_GROUND);
    }
    seed_row[static_cast<size_t>(blockIdx.y) * kp.W + c] = r;
    eseed[static_cast<size_t>(blockIdx.y) * kp.W + c] = (accepted && !valid) ? 1 : 0;
}

__device__ __forceinline__ bool is_ground(const int* par, int p) { return (par[uf_find(par, p)] & FLAG_GROUND) != 0; }

// The Labeler's image L: non-empty pixels of flagged components; accepted empty seeds; empty
// pixels 4-adjacent to a ground pixel u with |angle(u) - angle(empty)| <= K (labelled by the BFS but not expanded,
// label.cpp:67-74).
__global__ void __launch_bounds__(256) k_label(const uint16_t* __restrict__ range, const uint16_t* __restrict__ angle,
                                               const int* __restrict__ parent, const int* __restrict__ seed_row,
                                               const uint8_t* __restrict__ eseed, uint8_t* __restrict__ label, KParams kp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kp.P) return;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    const uint16_t* R = range + base;
    const uint16_t* A = angle + base;
    const int* par = parent + base;
    unsigned L = 0;
    if (R[p] != 0) {
        L = is_ground(par, p);
    } else {
        const int r = p / kp.W, c = p - r * kp.W;
        const size_t cb = static_cast<size_t>(blockIdx.y) * kp.W + c;
        if (seed_row[cb] == r && eseed[cb]) L = 1;
        const int a = A[p];
        if (!L && r > 0) {
            const int q = p - kp.W;
            if (R[q] != 0 && abs(a - static_cast<int>(A[q])) <= kp.k_ground) L = is_ground(par, q);
        }
        if (!L && r < kp.H - 1) {
            const int q = p + kp.W;
            if (R[q] != 0 && abs(a - static_cast<int>(A[q])) <= kp.k_ground) L = is_ground(par, q);
        }
        if (!L && kp.W > 1) {
            int q = (c > 0) ? p - 1 : p + kp.W - 1;
            if (R[q] != 0 && abs(a - static_cast<int>(A[q])) <= kp.k_ground) L = is_ground(par, q);
            if (!L) {
                q = (c < kp.W - 1) ? p + 1 : p - (kp.W - 1);
                if (R[q] != 0 && abs(a - static_cast<int>(A[q])) <= kp.k_ground) L = is_ground(par, q);
            }
        }
    }
    label[base + p] = static_cast<uint8_t>(L);
}

// Label::LabelOneComponent on an existing label image: flag the seed's component, then write `label` into every
// unlabelled pixel the BFS would reach.  Single frame.
__global__ void k_one_seed(const uint16_t* __restrict__ range, const uint8_t* __restrict__ label_image, int* __restrict__ parent,
                           int seed, int* __restrict__ seed_state) {
    // seed_state: 0 = seed already labelled (nothing to do), 1 = empty seed (labelled, not expanded), 2 = component flagged
    if (label_image[seed] != 0) { *seed_state = 0; return; }
    if (range[seed] == 0) { *seed_state = 1; return; }
    atomicOr(parent + uf_find(parent, seed), FLAG_GROUND);
    *seed_state = 2;
}

__global__ void __launch_bounds__(256) k_one_label(const uint16_t* __restrict__ R, const uint16_t* __restrict__ A,
                                                   const int* __restrict__ par, const int* __restrict__ seed_state, int seed,
                                                   int k_edge, int label, uint8_t* __restrict__ label_image, KParams kp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kp.P) return;
    const int st = *seed_state;
    if (st == 0 || label_image[p] != 0) return;
    if (st == 1) {
        if (p == seed) label_image[p] = static_cast<uint8_t>(label);
        return;
    }
    bool hit = false;
    if (R[p] != 0) {
        hit = is_ground(par, p);
    } else {
        // an unlabelled empty pixel next to a reached pixel u (u is non-empty and was unlabelled before the call)
        const int r = p / kp.W, c = p - r * kp.W, a = A[p];
        int q[4];
        q[0] = r > 0 ? p - kp.W : -1;
        q[1] = r < kp.H - 1 ? p + kp.W : -1;
        q[2] = kp.W > 1 ? (c > 0 ? p - 1 : p + kp.W - 1) : -1;
        q[3] = kp.W > 1 ? (c < kp.W - 1 ? p + 1 : p - (kp.W - 1)) : -1;
        for (int k = 0; k < 4 && !hit; ++k) {
            const int u = q[k];
            if (u < 0 || R[u] == 0 || abs(a - static_cast<int>(A[u])) > k_edge) continue;
            // u must be part of the flagged forest: blocked (pre-labelled) pixels are singletons that were never united
            hit = (par[uf_find(par, u)] & FLAG_GROUND) != 0;
        }
    }
    // NOTE: pixels are tested against the label image as it was BEFORE this launch for non-empty pixels only through
    // `par` (built before); reading label_image[p] != 0 above only skips pixels this launch never writes twice.
    if (hit) label_image[p] = static_cast<uint8_t>(label);
}

// All-components mode: canonical label = root = smallest flat index of the component, -1 for empty pixels.
__global__ void __launch_bounds__(256) k_components(const uint16_t* __restrict__ image, int* __restrict__ parent,
                                                    int32_t* __restrict__ labels, Counters* cnt, KParams kp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= kp.P) return;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    int out = -1;
    if (image[base + p] != 0) {
        out = uf_find(parent + base, p);
        if (out == p) atomicAdd(&cnt->components, 1ull);
    }
    labels[base + p] = out;
}

// ------------------------------------------------------------------------------------------------
// K3b — the Labeler as an all-components CCL on whole batches (north-star kernel 4, BASELINE configs C3 / C5):
// Label::LabelOneComponent (Labeler/label.cpp:48-104) driven over every unlabelled non-empty pixel = connected components
// of the non-empty pixels under |angle difference| <= K_beta, 4-neighbours with column wrap (Label::WrapCols, :33-45);
// canonical label = smallest flat index of the component, -1 for empty pixels.
//   k_ccl_tile    one CTA per (H x TW tile, frame): the tile's angles in shared memory, vertical runs labelled by a column
//                 walk (no atomics), one union per pair of horizontally linked runs (the link of row r is skipped when row
//                 r-1 joins the same two runs), roots = smallest index; writes the tile-local result as flat indices
//   k_ccl_seams   unions across the tile seams (and the wrap seam W-1 | 0) on the label image itself
//   k_ccl_flatten every pixel takes its root's index; counts the components of each frame
// The label image is its own union-find forest between the kernels: labels[p] = flat index of p's parent, roots point to
// themselves and are the smallest index of their tree.
// ------------------------------------------------------------------------------------------------
// find with path halving.  Parents only ever move to smaller indices of the same tree (roots are hung with atomicMin, the
// halving step uses atomicMin too), so concurrent finds and unions cannot lose a link; without the halving the "smaller index
// wins" rule builds chains as long as a component is wide and every find walks them.
__device__ __forceinline__ int ufs_find(int* par, int i) {
    for (;;) {
        const int p = *reinterpret_cast<volatile int*>(par + i);
        if (p == i) return i;
        const int gp = *reinterpret_cast<volatile int*>(par + p);
        if (gp == p) return p;
        atomicMin(par + i, gp);
        i = gp;
    }
}
__device__ __forceinline__ void ufs_unite(int* par, int a, int b) {  // shared or global forest, the smaller index becomes the root
    for (;;) {
        a = ufs_find(par, a);
        b = ufs_find(par, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(par + a, b);
        if (old == a) return;
        a = old;
    }
}

constexpr unsigned short CCL_EMPTY = 0xffffu;

// TW = 1 << TWL columns per tile (a power of two: row / column of a tile-local index are a shift and a mask)
__global__ void __launch_bounds__(256) k_ccl_tile(const uint16_t* __restrict__ image, const uint16_t* __restrict__ angle, int k_edge, int TWL,
                                                  int32_t* __restrict__ labels, KParams kp) {
    extern __shared__ __align__(16) unsigned char ccl_smem[];
    const int H = kp.H, W = kp.W, TW = 1 << TWL, CM = TW - 1;
    const int c0 = blockIdx.x << TWL, tw = min(TW, W - c0);
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    int* par = reinterpret_cast<int*>(ccl_smem);                               // [H][TW] tile-local row-major index of the parent
    unsigned short* ang = reinterpret_cast<unsigned short*>(par + H * TW);     // [H][TW] angle, CCL_EMPTY for empty pixels
    const int n = H * TW;
    // both images are read unconditionally and four pixels ahead: the tile load is the one phase that waits for HBM / L2
    for (int t0 = threadIdx.x; t0 < n; t0 += 4 * blockDim.x) {
        unsigned img[4], an[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = t0 + j * blockDim.x;
            const int r = t >> TWL, c = t & CM;
            img[j] = 0;
            an[j] = 0;
            if (t < n && c < tw) {
                const size_t g = base + static_cast<size_t>(r) * W + c0 + c;
                img[j] = __ldg(image + g);
                an[j] = __ldg(angle + g);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = t0 + j * blockDim.x;
            if (t < n) ang[t] = img[j] ? static_cast<unsigned short>(min(an[j], 0xfffeu)) : CCL_EMPTY;
        }
    }
    __syncthreads();
    auto linked = [&](int t, int u) -> bool {
        const int a = ang[t], b = ang[u];
        return a != CCL_EMPTY && b != CCL_EMPTY && abs(a - b) <= k_edge;
    };
    // vertical runs: the top pixel of a run is its root.  Every column is walked in segments of SEG rows by different threads (all
    // threads of the CTA busy); a run that continues across a segment boundary is joined below like any other pair of runs.
    const int n_seg = max(1, min(static_cast<int>(blockDim.x) >> TWL, H >> 3));
    const int SEG = (H + n_seg - 1) / n_seg;
    for (int job = threadIdx.x; job < (n_seg << TWL); job += blockDim.x) {
        const int c = job & CM, r0 = (job >> TWL) * SEG, r1 = min(r0 + SEG, H);
        if (c < tw) {
            int top = r0 * TW + c;
            for (int r = r0; r < r1; ++r) {
                const int t = r * TW + c;
                if (r == r0 || !linked(t - TW, t)) top = t;
                par[t] = top;
            }
        }
    }
    __syncthreads();
    for (int job = threadIdx.x; job < (n_seg << TWL); job += blockDim.x) {
        const int c = job & CM, r0 = (job >> TWL) * SEG;
        const bool todo = c < tw && r0 > 0 && r0 < H && linked((r0 - 1) * TW + c, r0 * TW + c);
        if (todo) ufs_unite(par, (r0 - 1) * TW + c, r0 * TW + c);
        __syncwarp();
    }
    __syncthreads();
    // horizontal links: one union per pair of runs.  The loops keep whole warps together (no early `continue`, a __syncwarp per
    // iteration): with lanes left to drift apart inside the union's retry loop, every lane subset walks the rest of the loop on
    // its own and the loop header is issued an order of magnitude more often (measured: 363 instead of 32 issues per warp).
    for (int t0 = 0; t0 < n; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        const int r = t >> TWL, c = t & CM;
        bool todo = t < n && c + 1 < tw && linked(t, t + 1);
        if (todo && r > 0 && linked(t - TW, t - TW + 1) && linked(t - TW, t) && linked(t - TW + 1, t + 1)) todo = false;  // row r-1 joins the same runs
        if (todo) ufs_unite(par, t, t + 1);
        __syncwarp();
    }
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        const int r = t >> TWL, c = t & CM;
        const bool in = t < n && c < tw;
        int out = -1;
        if (in && ang[t] != CCL_EMPTY) {
            const int root = ufs_find(par, t);
            out = (root >> TWL) * W + c0 + (root & CM);
        }
        __syncwarp();
        if (in) labels[base + static_cast<size_t>(r) * W + c0 + c] = out;
    }
}

__global__ void __launch_bounds__(256) k_ccl_seams(const uint16_t* __restrict__ image, const uint16_t* __restrict__ angle, int k_edge, int TW,
                                                   int n_tiles, int32_t* __restrict__ labels, KParams kp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kp.H * n_tiles) return;
    const int r = i / n_tiles, b = i - r * n_tiles;
    // seam b: last column of tile b | first column of tile b + 1; the last seam is the wrap W-1 | 0
    const int cl = min((b + 1) * TW, kp.W) - 1;
    const int cr = (cl + 1 == kp.W) ? 0 : cl + 1;
    if (cl == cr) return;  // one-column image
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    const int pl = r * kp.W + cl, pr = r * kp.W + cr;
    if (image[base + pl] == 0 || image[base + pr] == 0) return;
    if (abs(static_cast<int>(angle[base + pl]) - static_cast<int>(angle[base + pr])) > k_edge) return;
    ufs_unite(labels + base, pl, pr);
}

__global__ void __launch_bounds__(256) k_ccl_flatten(int32_t* __restrict__ labels, unsigned* __restrict__ n_components, Counters* cnt, KParams kp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool root = false;
    if (p < kp.P) {
        int32_t* L = labels + static_cast<size_t>(blockIdx.y) * kp.P;
        if (L[p] >= 0) {
            const int rt = ufs_find(L, p);
            L[p] = rt;
            root = rt == p;
        }
    }
    const int n_roots = __syncthreads_count(root);  // one atomic per block: thousands of roots per frame would queue on one L2 address
    if (threadIdx.x == 0 && n_roots) {
        if (n_components) atomicAdd(n_components + blockIdx.y, static_cast<unsigned>(n_roots));
        if (cnt) atomicAdd(&cnt->components, static_cast<unsigned long long>(n_roots));
    }
}

// ------------------------------------------------------------------------------------------------
// N4 — the FPGA wire formats of AlfaNode (src/alfa_node.cpp:105-271) and Alfa_GS::CreateColoredAngleImage (src/alfa_gs.cpp:768-829)
// ------------------------------------------------------------------------------------------------
// AlfaNode::store_pointcloud_hardware (alfa_node.cpp:105-131): one 64-bit word per point = four u16 {elevation, azimuth in
// centi-degrees, range in centimetres, 0}.  The reference's arithmetic is float where the overloads make it float:
// std::hypot / std::atan2 / std::sqrt on float arguments are hypotf / atan2f / sqrtf, the degree factor 180.0f / M_PI is a
// double, the product is cast to float before the * 100 (alfa_node.cpp:115-120).  atan2f is glibc's, restated in gs_math.cuh;
// glibc's hypotf is sqrt((double)x*x + (double)y*y) rounded to float.
__device__ __forceinline__ unsigned short f2u16_like_x86(float v) { return static_cast<unsigned short>(__float2int_rz(v)); }  // cvttss2si, low 16 bits

__global__ void __launch_bounds__(256) k_fpga_store(const unsigned char* __restrict__ points, size_t n, size_t stride,
                                                    unsigned long long* __restrict__ words) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = reinterpret_cast<const float*>(points + i * stride);
    const float x = q[0], y = q[1], z = q[2];
    const double rad2deg = static_cast<double>(180.0f) / 3.14159265358979323846;  // 180.0f / M_PI
    const double xd = x, yd = y;
    const float h = __double2float_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(xd, xd), __dmul_rn(yd, yd))));  // hypotf
    const float el = gs_atan2f(z, h);
    const short elevation = static_cast<short>(__float2int_rz(__fmul_rn(__double2float_rn(__dmul_rn(static_cast<double>(el), rad2deg)), 100.0f)));  // :115
    const float a = gs_atan2f(y, x);                                                                    // :117
    const double az = (y >= 0.0f) ? static_cast<double>(a) : __dadd_rn(static_cast<double>(a), 6.283185307179586476925286766559);
    const unsigned short azimuth = f2u16_like_x86(__fmul_rn(__double2float_rn(__dmul_rn(az, rad2deg)), 100.0f));  // :118
    const float r = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    const unsigned short range = f2u16_like_x86(__fmul_rn(r, 100.0f));                                  // :119
    words[i] = static_cast<unsigned long long>(static_cast<unsigned short>(elevation)) | (static_cast<unsigned long long>(azimuth) << 16) |
               (static_cast<unsigned long long>(range) << 32);
}

// AlfaNode::read_hardware_pointcloud (alfa_node.cpp:157-194): four u16 pixels per word, column by column (64 rows per column),
// into the 64 x 1800 image the reference hard-codes (:162)
__global__ void __launch_bounds__(256) k_fpga_read_range(const unsigned long long* __restrict__ words, unsigned n_pixels, uint16_t* __restrict__ image) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pixels) return;
    const unsigned v = static_cast<unsigned>(words[k >> 2] >> (16 * (k & 3))) & 0xffffu;
    image[(k & 63u) * 1800u + (k >> 6)] = static_cast<uint16_t>(v);
}

// AlfaNode::read_hardware_filtered_angle_image (alfa_node.cpp:196-271): 10-bit fields, six per word, every 43rd word four
// (bursts of 256 pixels in 43 words); pixels again column by column, 64 rows per column, into a rows x cols image
__global__ void __launch_bounds__(256) k_fpga_read_angle(const unsigned long long* __restrict__ words, unsigned n_pixels, unsigned cols,
                                                         uint16_t* __restrict__ image) {
    const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_pixels) return;
    const unsigned burst = n >> 8, within = n & 255u;
    const unsigned word = burst * 43u + (within < 252u ? within / 6u : 42u);
    const unsigned field = within < 252u ? within % 6u : within - 252u;
    image[(n & 63u) * cols + (n >> 6)] = static_cast<uint16_t>((words[word] >> (10u * field)) & 0x3ffu);
}

// Alfa_GS::CreateColoredAngleImage (alfa_gs.cpp:768-829): the scale is 255 / (the largest angle seen SO FAR in raster order,
// :781-793), i.e. an inclusive prefix maximum over the flat index; one block walks the image in chunks with a running carry.
__global__ void __launch_bounds__(1024) k_colored_angle(const uint16_t* __restrict__ angle, unsigned n_pixels, unsigned char* __restrict__ rgb) {
    __shared__ unsigned warp_max[32];
    __shared__ unsigned carry_s;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (unsigned base = 0; base < n_pixels; base += 1024) {
        const unsigned i = base + threadIdx.x;
        const unsigned a = i < n_pixels ? angle[i] : 0u;
        unsigned m = a;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) m = max(m, __shfl_up_sync(0xffffffffu, m, d) * (lane >= static_cast<unsigned>(d) ? 1u : 0u));
        if (lane == 31) warp_max[warp] = m;
        __syncthreads();
        unsigned before = carry_s;
        for (unsigned w = 0; w < warp; ++w) before = max(before, warp_max[w]);
        const unsigned mx = max(m, before);  // max_angle_image after this pixel's update (:781-782)
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = mx;
        if (i < n_pixels) {
            // ushort angle = (255.0 / max_angle_image) * angle (:793); 255.0 / 0 * 0 is NaN, which x86 converts to 0
            const unsigned s = mx ? static_cast<unsigned>(__double2int_rz(__dmul_rn(__ddiv_rn(255.0, static_cast<double>(mx)), static_cast<double>(a)))) & 0xffffu : 0u;
            unsigned char c0, c1, c2;
            if (s == 0) { c0 = 0; c1 = 0; c2 = 0; }
            else if (s >= 128) { c0 = static_cast<unsigned char>((s - 128) * 2); c1 = static_cast<unsigned char>(255 - ((s - 128) * 2)); c2 = 0; }
            else { c0 = 0; c1 = static_cast<unsigned char>(s * 2); c2 = static_cast<unsigned char>(255 - 2 * s); }
            rgb[3 * i] = c0;
            rgb[3 * i + 1] = c1;
            rgb[3 * i + 2] = c2;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K4 — dilation taps + erase, then ordered compaction + back-projection.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned block_sum_256(unsigned v, unsigned* smem) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    unsigned tot = 0;
    for (int w = 0; w < 8; ++w) tot += smem[w];
    __syncthreads();
    return tot;
}

// One block per image row.  dilated(r,c) = OR_k L(r + off_k, c) — the effective structuring element of
// cv::dilate with GetUniformKernel's bytes; res = dilated ? 0 : range (alfa_gs.cpp:751-764).
__global__ void __launch_bounds__(256) k_erase(const uint16_t* __restrict__ range, const uint8_t* __restrict__ label,
                                               uint16_t* __restrict__ res, unsigned* __restrict__ rowcount, KParams kp) {
    __shared__ unsigned smem[8];
    const int r = blockIdx.x;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    const int r0 = r + kp.dil0, r1 = r + kp.dil1;
    const bool ok0 = r0 >= 0 && r0 < kp.H, ok1 = kp.n_dil > 1 && r1 >= 0 && r1 < kp.H;
    unsigned count = 0;
    for (int c = threadIdx.x; c < kp.W; c += 256) {
        unsigned dil = 0;
        if (ok0) dil |= label[base + r0 * kp.W + c];
        if (ok1) dil |= label[base + r1 * kp.W + c];
        const unsigned v = range[base + r * kp.W + c];
        const unsigned o = (kp.erase_ground ? (dil == 0) : (dil != 0)) ? v : 0u;
        res[base + r * kp.W + c] = static_cast<uint16_t>(o);
        count += (o != 0);
    }
    const unsigned tot = block_sum_256(count, smem);
    if (threadIdx.x == 0) rowcount[static_cast<size_t>(blockIdx.y) * kp.H + r] = tot;
}

__global__ void __launch_bounds__(256) k_rowcount(const uint16_t* __restrict__ image, unsigned* __restrict__ rowcount, KParams kp) {
    __shared__ unsigned smem[8];
    const int r = blockIdx.x;
    const size_t base = static_cast<size_t>(blockIdx.y) * kp.P;
    unsigned count = 0;
    for (int c = threadIdx.x; c < kp.W; c += 256) count += image[base + r * kp.W + c] != 0;
    const unsigned tot = block_sum_256(count, smem);
    if (threadIdx.x == 0) rowcount[static_cast<size_t>(blockIdx.y) * kp.H + r] = tot;
}

// One block per image row: output slot = (non-zero pixels in the rows above) + rank inside the row, which is exactly the
// row-major push_back order of CameraCb (alfa_gs.cpp:1023-1058).  cos/sin come from host double LUTs, so x,y,z are the
// reference's doubles rounded to float.
template <bool RANGE_AS_DOUBLE>
__global__ void __launch_bounds__(256) k_backproject(const uint16_t* __restrict__ image, const unsigned* __restrict__ rowcount,
                                                     float4* __restrict__ xyz, unsigned* __restrict__ counts, KParams kp, Luts lut) {
    __shared__ unsigned warp_tot[8];
    __shared__ unsigned s_base;
    const int r = blockIdx.x;
    const int f = blockIdx.y;
    const size_t base = static_cast<size_t>(f) * kp.P;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // rows above: H <= 1024 counters, summed by the first warp
    if (warp == 0) {
        unsigned s = 0;
        for (int q = lane; q < r; q += 32) s += rowcount[static_cast<size_t>(f) * kp.H + q];
        s = __reduce_add_sync(0xffffffffu, s);
        if (lane == 0) s_base = s;
    }
    __syncthreads();
    unsigned running = s_base;
    const double ca = lut.cos_alt[r], sa = lut.sin_alt[r];
    float4* out = xyz + base;  // slab of frame f: up to P records
    for (int c0 = 0; c0 < kp.W; c0 += 256) {
        const int c = c0 + threadIdx.x;
        const unsigned enc = (c < kp.W) ? image[base + r * kp.W + c] : 0u;
        const unsigned ballot = __ballot_sync(0xffffffffu, enc != 0);
        if (lane == 0) warp_tot[warp] = __popc(ballot);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < 8; ++w) {
            const unsigned t = warp_tot[w];
            before += (w < static_cast<int>(warp)) ? t : 0u;
            total += t;
        }
        if (enc != 0 && xyz) {  // xyz == nullptr: the caller wants the counts only
            const unsigned slot = running + before + __popc(ballot & ((1u << lane) - 1u));
            const double rng_d = decode_range(kp, enc);
            const double rng = RANGE_AS_DOUBLE ? rng_d : static_cast<double>(__double2float_rn(rng_d));  // alfa_gs.cpp:1027,1045 vs range2cloud_node.cpp:76
            const double caz = lut.cos_az[c], saz = lut.sin_az[c];
            float4 o;
            o.x = __double2float_rn(__dmul_rn(__dmul_rn(ca, caz), rng));  // :1055
            o.y = __double2float_rn(__dmul_rn(__dmul_rn(ca, saz), rng));  // :1056
            o.z = __double2float_rn(__dmul_rn(sa, rng));                  // :1057
            o.w = 1.0f;
            out[slot] = o;
        }
        running += total;
        __syncthreads();
    }
    if (r == kp.H - 1 && threadIdx.x == 0) counts[f] = running;
}

// ------------------------------------------------------------------------------------------------
// N1 — wire formats either side of the path.  sensor_msgs/PointCloud2 carries points as `point_step`-byte records with
// float32 fields at arbitrary offsets (pcl::fromROSMsg, alfa_node.cpp:311); the pipeline wants packed 16-byte XYZI.
// The cloud going out (pcl::toROSMsg, alfa_node.cpp:35-41) is PointXYZ (= our {x,y,z,1} records) or PointXYZRGB (32 bytes).
// Both are pure bandwidth kernels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_f32(const unsigned char* p, bool aligned, bool swap) {
    unsigned v;
    if (aligned) v = *reinterpret_cast<const unsigned*>(p);
    else v = static_cast<unsigned>(p[0]) | (static_cast<unsigned>(p[1]) << 8) | (static_cast<unsigned>(p[2]) << 16) | (static_cast<unsigned>(p[3]) << 24);
    if (swap) v = __byte_perm(v, 0, 0x0123);
    return __uint_as_float(v);
}

__global__ void __launch_bounds__(256) k_unpack_points(const unsigned char* __restrict__ data, size_t n, unsigned step, unsigned ox, unsigned oy,
                                                       unsigned oz, int oi, int big_endian, float4* __restrict__ out) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char* p = data + i * step;
    const bool aligned = ((reinterpret_cast<uintptr_t>(data) | step | ox | oy | oz | (oi >= 0 ? static_cast<unsigned>(oi) : 0u)) & 3u) == 0;
    float4 q;
    q.x = load_f32(p + ox, aligned, big_endian != 0);
    q.y = load_f32(p + oy, aligned, big_endian != 0);
    q.z = load_f32(p + oz, aligned, big_endian != 0);
    q.w = oi >= 0 ? load_f32(p + oi, aligned, big_endian != 0) : 0.0f;
    out[i] = q;
}

// {x, y, z, 1} -> pcl::PointXYZRGB: x y z 1.0 | b g r a = 0 0 0 255 | 12 bytes of padding   (CameraCb's points, alfa_gs.cpp:1055-1058)
__global__ void __launch_bounds__(256) k_pack_xyzrgb(const float4* __restrict__ xyz4, unsigned n, uint4* __restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q = xyz4[i];
    out[2 * i] = make_uint4(__float_as_uint(q.x), __float_as_uint(q.y), __float_as_uint(q.z), __float_as_uint(1.0f));
    out[2 * i + 1] = make_uint4(0xff000000u, 0u, 0u, 0u);
}


// ------------------------------------------------------------------------------------------------
// N3 — evaluation (Alfa_GS::CreateLabeledRangeImage alfa_gs.cpp:945-995, CheckNumberOfDetectedRIdGnd :831-943)
// ------------------------------------------------------------------------------------------------
// Labelled range image: 1 everywhere, 0 where a kept point lands while the reference's `gnd_flag` is up.  The flag is
// raised by every point with a ground colour (:961-965) BEFORE the range / altitude gates and lowered only at the end of
// a kept point's iteration (:992), so it sticks across skipped points: a kept point is marked iff itself or any point
// since the previous kept point has a ground colour.  Pixels are only ever set to 0, hence the result does not depend on
// the order of the stores; the sticky flag is a segmented OR over the input order, done in three passes: per-point
// (kept, ground) bits, per-chunk summaries, look-back.
constexpr int EV_CHUNK = 256;
__device__ __forceinline__ bool ground_colour(unsigned rgba) {
    // PCL PointXYZRGB: bytes b, g, r, a at offset 16 (alfa_gs.cpp:961)
    const unsigned b = rgba & 0xffu, g = (rgba >> 8) & 0xffu, r = (rgba >> 16) & 0xffu;
    return (r == 75 && g == 0 && b == 75) || (r == 255 && g == 0 && b == 255) || (r == 255 && g == 150 && b == 255);
}
__global__ void __launch_bounds__(EV_CHUNK) k_label_flags(const unsigned char* __restrict__ points, size_t n, KParams kp, uint8_t* __restrict__ flags,
                                                           int* __restrict__ flat_out) {
    const size_t i = static_cast<size_t>(blockIdx.x) * EV_CHUNK + threadIdx.x;
    if (i >= n) return;
    const float* q = reinterpret_cast<const float*>(points + i * 32);
    const unsigned rgba = *reinterpret_cast<const unsigned*>(points + i * 32 + 16);
    const Projected pr = project_point(kp, q[0], q[1], q[2]);
    flags[i] = static_cast<uint8_t>((pr.flat >= 0 ? 1 : 0) | (ground_colour(rgba) ? 2 : 0));
    flat_out[i] = pr.flat;
}
// one thread per chunk: (a kept point exists) | (ground colour after the last kept point) << 1 | (ground colour anywhere) << 2
__global__ void __launch_bounds__(128) k_label_chunks(const uint8_t* __restrict__ flags, size_t n, uint8_t* __restrict__ summary, int n_chunks) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const size_t beg = static_cast<size_t>(c) * EV_CHUNK, end = min(beg + EV_CHUNK, n);
    unsigned kept = 0, tail = 0, all = 0;
    for (size_t i = beg; i < end; ++i) {
        const unsigned f = flags[i];
        if (f & 2u) { tail = 1; all = 1; }
        if (f & 1u) { kept = 1; tail = 0; }
    }
    summary[c] = static_cast<uint8_t>(kept | (tail << 1) | (all << 2));
}
__global__ void __launch_bounds__(EV_CHUNK) k_label_mark(const uint8_t* __restrict__ flags, const int* __restrict__ flat, const uint8_t* __restrict__ summary,
                                                          size_t n, uint16_t* __restrict__ labeled) {
    const size_t i = static_cast<size_t>(blockIdx.x) * EV_CHUNK + threadIdx.x;
    if (i >= n) return;
    const unsigned f = flags[i];
    if (!(f & 1u)) return;  // skipped points store nothing
    bool up = (f & 2u) != 0;
    // look back to the previous kept point: inside the chunk point by point, then over chunk summaries
    const size_t chunk_beg = static_cast<size_t>(blockIdx.x) * EV_CHUNK;
    bool found_kept = false;
    for (size_t j = i; !up && j > chunk_beg;) {
        --j;
        const unsigned g = flags[j];
        if (g & 1u) { found_kept = true; break; }
        if (g & 2u) up = true;
    }
    if (!up && !found_kept) {
        for (int c = static_cast<int>(blockIdx.x) - 1; c >= 0; --c) {
            const unsigned sm = summary[c];
            if (sm & 1u) { up = (sm & 2u) != 0; break; }
            if (sm & 4u) { up = true; break; }
        }
    }
    if (up) labeled[flat[i]] = 0;  // alfa_gs.cpp:988-991
}
__global__ void __launch_bounds__(256) k_fill_u16(uint16_t* __restrict__ p, size_t n, uint16_t v) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// CheckNumberOfDetectedRIdGnd: the nine per-frame counters of alfa_gs.cpp:838-870 in one pass over the three images
__global__ void __launch_bounds__(256) k_evaluate(const uint16_t* __restrict__ og, const uint16_t* __restrict__ seg, const uint16_t* __restrict__ lab, size_t P,
                                                  unsigned long long* __restrict__ out9) {
    unsigned c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) c[k] = 0;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < P; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const unsigned o = og[i], s = seg[i], l = lab[i];
        c[0] += 1;                                 // points_cntr
        if (l == 0) c[1] += 1;                     // labeled_gnd_cntr          :840
        else if (o != 0) c[2] += 1;                // labeled_obst_cntr         :843-845
        if (s == 0 && o != 0) c[3] += 1;           // seg_gnd_cntr              :847
        if (s == l && o != 0) c[4] += 1;           // true_positive             :850
        if (s == 0 && l != 0 && o != 0) c[5] += 1; // false_positive            :855
        if (s != 0 && l != 0) c[6] += 1;           // true_negative             :858
        if (s != 0 && l == 0) c[7] += 1;           // false_negative            :861
        if (s != 0 && l != 0) c[8] += 1;           // true_detected_obst_cntr   :864
    }
    __shared__ unsigned sm[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        unsigned v = c[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += sm[threadIdx.x][w];
        if (t) atomicAdd(out9 + threadIdx.x, t);
    }
}

// ------------------------------------------------------------------------------------------------
// host batch path: the clouds of a wave, one slab of P records per frame, are packed back to back on the device so that
// they leave with ONE device-to-host copy (a copy per frame costs ~10 us of DMA set-up each and stalls the pipeline)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_count_prefix(const unsigned* __restrict__ counts, int n, unsigned long long* __restrict__ prefix) {
    // exclusive prefix sum of up to a few thousand frame counts by one block
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n ? counts[i] : 0ull;
        unsigned long long incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long before = carry_s;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) before += warp_sum[w];
        if (i < n) prefix[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) prefix[n] = carry_s;
}
__global__ void __launch_bounds__(256) k_pack_clouds(const float4* __restrict__ slabs, const unsigned* __restrict__ counts,
                                                     const unsigned long long* __restrict__ prefix, size_t P, float4* __restrict__ packed) {
    const int f = blockIdx.y;
    const unsigned n = counts[f];
    const float4* src = slabs + static_cast<size_t>(f) * P;
    float4* dst = packed + prefix[f];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

inline dim3 pixel_grid(const KParams& kp, int frames) { return dim3((kp.P + 255) / 256, frames); }

}  // namespace

#define GS_COUNT(L) (++*(L).n_launches)

void launch_unpack_points(const Launch& L, const void* data, size_t n, unsigned step, unsigned ox, unsigned oy, unsigned oz, int oi, int big_endian,
                          float* xyzi_out) {
    if (n == 0) return;
    k_unpack_points<<<static_cast<unsigned>((n + 255) / 256), 256, 0, L.stream>>>(static_cast<const unsigned char*>(data), n, step, ox, oy, oz, oi,
                                                                                 big_endian, reinterpret_cast<float4*>(xyzi_out));
    GS_COUNT(L);
}

void launch_pack_xyzrgb(const Launch& L, const float* xyz4, unsigned n, void* out) {
    if (n == 0) return;
    k_pack_xyzrgb<<<(n + 255) / 256, 256, 0, L.stream>>>(reinterpret_cast<const float4*>(xyz4), n, static_cast<uint4*>(out));
    GS_COUNT(L);
}

void launch_project(const Launch& L, const KParams& kp, const void* points, size_t stride, const long long* d_offsets, int frames,
                    long long max_points, unsigned long long* keys, Counters* cnt, bool stats) {
    if (max_points <= 0 || frames <= 0) return;
    const dim3 grid(static_cast<unsigned>((max_points + 255) / 256), frames);
    const bool s16 = stride == 16 && (reinterpret_cast<uintptr_t>(points) & 15) == 0;
    if (s16) k_project<true><<<grid, 256, 0, L.stream>>>(static_cast<const char*>(points), stride, d_offsets, kp, keys, cnt, stats);
    else k_project<false><<<grid, 256, 0, L.stream>>>(static_cast<const char*>(points), stride, d_offsets, kp, keys, cnt, stats);
    GS_COUNT(L);
}

void launch_finalize(const Launch& L, const KParams& kp, int frames, unsigned long long* keys, uint16_t* range) {
    k_finalize<<<pixel_grid(kp, frames), 256, 0, L.stream>>>(keys, range, kp);
    GS_COUNT(L);
}

void launch_angle(const Launch& L, const KParams& kp, const Luts& lut, int frames, const uint16_t* range, uint16_t* angle, Counters* cnt,
                  bool stats) {
    k_angle<<<pixel_grid(kp, frames), 256, 0, L.stream>>>(range, angle, kp, lut, cnt, stats);
    GS_COUNT(L);
}

void launch_sg(const Launch& L, const KParams& kp, const SgTaps& taps, int frames, const uint16_t* in, uint16_t* out) {
    k_sg<<<pixel_grid(kp, frames), 256, 0, L.stream>>>(in, out, kp, taps);
    GS_COUNT(L);
}

void launch_repair(const Launch& L, const KParams& kp, int frames, const uint16_t* in, uint16_t* out, int step, float thr) {
    k_repair<<<dim3((kp.W + 127) / 128, frames), 128, 0, L.stream>>>(in, out, kp, step, thr);
    GS_COUNT(L);
}

void launch_ccl(const Launch& L, const KParams& kp, int frames, const uint16_t* range, const uint16_t* angle, const uint8_t* blocked,
                int k_edge, int* parent) {
    k_ccl_init<<<pixel_grid(kp, frames), 256, 0, L.stream>>>(parent, kp);
    GS_COUNT(L);
    k_ccl_union<<<pixel_grid(kp, frames), 256, 0, L.stream>>>(range, angle, blocked, k_edge, parent, kp);
    GS_COUNT(L);
}

void launch_ground(const Launch& L, const KParams& kp, int frames, const uint16_t* range, const uint16_t* angle, int* parent,
                   int* seed_row, uint8_t* eseed, uint8_t* label) {
    k_seeds<<<dim3((kp.W + 127) / 128, frames), 128, 0, L.stream>>>(range, angle, parent, seed_row, eseed, kp);
    GS_COUNT(L);
    k_label<<<pixel_grid(kp, frames), 256, 0, L.stream>>>(range, angle, parent, seed_row, eseed, label, kp);
    GS_COUNT(L);
}

void launch_one_component(const Launch& L, const KParams& kp, const uint16_t* range, const uint16_t* angle, int k_edge, int* parent,
                          int row, int col, int label, uint8_t* label_image) {
    // parent was built by launch_ccl with blocked = label_image; seed_row scratch doubles as the seed state word
    int* seed_state = parent + kp.P;  // caller guarantees one spare int after the forest
    const int seed = row * kp.W + col;
    k_one_seed<<<1, 1, 0, L.stream>>>(range, label_image, parent, seed, seed_state);
    GS_COUNT(L);
    k_one_label<<<pixel_grid(kp, 1), 256, 0, L.stream>>>(range, angle, parent, seed_state, seed, k_edge, label, label_image, kp);
    GS_COUNT(L);
}

void launch_components(const Launch& L, const KParams& kp, int frames, const uint16_t* image, int* parent, int32_t* labels,
                       Counters* cnt) {
    k_components<<<pixel_grid(kp, frames), 256, 0, L.stream>>>(image, parent, labels, cnt, kp);
    GS_COUNT(L);
}

void launch_fpga_store(const Launch& L, const void* points, size_t n, size_t stride, unsigned long long* words) {
    if (n == 0) return;
    k_fpga_store<<<static_cast<unsigned>((n + 255) / 256), 256, 0, L.stream>>>(static_cast<const unsigned char*>(points), n, stride, words);
    GS_COUNT(L);
}
void launch_fpga_read_range(const Launch& L, const unsigned long long* words, unsigned n_pixels, uint16_t* image64x1800) {
    if (n_pixels == 0) return;
    k_fpga_read_range<<<(n_pixels + 255) / 256, 256, 0, L.stream>>>(words, n_pixels, image64x1800);
    GS_COUNT(L);
}
void launch_fpga_read_angle(const Launch& L, const unsigned long long* words, unsigned n_pixels, unsigned cols, uint16_t* image) {
    if (n_pixels == 0) return;
    k_fpga_read_angle<<<(n_pixels + 255) / 256, 256, 0, L.stream>>>(words, n_pixels, cols, image);
    GS_COUNT(L);
}
void launch_colored_angle(const Launch& L, const uint16_t* angle, unsigned n_pixels, unsigned char* rgb) {
    if (n_pixels == 0) return;
    k_colored_angle<<<1, 1024, 0, L.stream>>>(angle, n_pixels, rgb);
    GS_COUNT(L);
}

int ccl_tile_width(const KParams& kp) {
    int tw = 512;
    while (tw > 8 && static_cast<long long>(tw) * kp.H > 8192) tw >>= 1;
    return tw;
}

void launch_components_batch(const Launch& L, const KParams& kp, int frames, const uint16_t* image, const uint16_t* angle, int k_edge,
                             int32_t* labels, unsigned* n_components, Counters* cnt) {
    const int TW = ccl_tile_width(kp);
    const int n_tiles = (kp.W + TW - 1) / TW;
    const size_t smem = static_cast<size_t>(kp.H) * TW * 6;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_ccl_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));  // per device
    if (n_components) cudaMemsetAsync(n_components, 0, static_cast<size_t>(frames) * sizeof(unsigned), L.stream);
    int TWL = 0;
    while ((1 << TWL) < TW) ++TWL;
    k_ccl_tile<<<dim3(n_tiles, frames), 256, smem, L.stream>>>(image, angle, k_edge, TWL, labels, kp);
    GS_COUNT(L);
    k_ccl_seams<<<dim3((kp.H * n_tiles + 255) / 256, frames), 256, 0, L.stream>>>(image, angle, k_edge, TW, n_tiles, labels, kp);
    GS_COUNT(L);
    k_ccl_flatten<<<pixel_grid(kp, frames), 256, 0, L.stream>>>(labels, n_components, cnt, kp);
    GS_COUNT(L);
}

void launch_erase(const Launch& L, const KParams& kp, int frames, const uint16_t* range, const uint8_t* label, uint16_t* res,
                  unsigned* rowcount) {
    k_erase<<<dim3(kp.H, frames), 256, 0, L.stream>>>(range, label, res, rowcount, kp);
    GS_COUNT(L);
}

void launch_rowcount(const Launch& L, const KParams& kp, int frames, const uint16_t* image, unsigned* rowcount) {
    k_rowcount<<<dim3(kp.H, frames), 256, 0, L.stream>>>(image, rowcount, kp);
    GS_COUNT(L);
}

void launch_backproject(const Launch& L, const KParams& kp, const Luts& lut, int frames, const uint16_t* image, const unsigned* rowcount,
                        float* xyz4, unsigned* counts, bool range_as_double) {
    if (range_as_double)
        k_backproject<true><<<dim3(kp.H, frames), 256, 0, L.stream>>>(image, rowcount, reinterpret_cast<float4*>(xyz4), counts, kp, lut);
    else
        k_backproject<false><<<dim3(kp.H, frames), 256, 0, L.stream>>>(image, rowcount, reinterpret_cast<float4*>(xyz4), counts, kp, lut);
    GS_COUNT(L);
}

void launch_label_image(const Launch& L, const KParams& kp, const void* points32, size_t n, uint8_t* flags, int* flat, uint8_t* summary,
                        uint16_t* labeled) {
    k_fill_u16<<<static_cast<unsigned>((kp.P + 255) / 256), 256, 0, L.stream>>>(labeled, static_cast<size_t>(kp.P), 1);  // Mat(n_beams_, n_cols_, CV_16UC1, 1)  :947
    ++*L.n_launches;
    if (n == 0) return;
    const int n_chunks = static_cast<int>((n + EV_CHUNK - 1) / EV_CHUNK);
    k_label_flags<<<n_chunks, EV_CHUNK, 0, L.stream>>>(static_cast<const unsigned char*>(points32), n, kp, flags, flat);
    k_label_chunks<<<(n_chunks + 127) / 128, 128, 0, L.stream>>>(flags, n, summary, n_chunks);
    k_label_mark<<<n_chunks, EV_CHUNK, 0, L.stream>>>(flags, flat, summary, n, labeled);
    *L.n_launches += 3;
}
size_t label_image_scratch_bytes(size_t n) {  // flags [n] u8 | summary [chunks] u8 (padded) | flat [n] i32
    const size_t chunks = (n + EV_CHUNK - 1) / EV_CHUNK;
    return ((n + 15) / 16 * 16) + ((chunks + 15) / 16 * 16) + n * 4;
}
void launch_evaluate(const Launch& L, const KParams& kp, const uint16_t* og, const uint16_t* seg, const uint16_t* lab, unsigned long long* out9) {
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>((kp.P + 255) / 256, 148 * 8));
    k_evaluate<<<blocks, 256, 0, L.stream>>>(og, seg, lab, static_cast<size_t>(kp.P), out9);
    ++*L.n_launches;
}

void launch_pack_clouds(const Launch& L, const float4* slabs, const unsigned* counts, int frames, size_t P, unsigned long long* prefix,
                        float4* packed) {
    if (frames <= 0) return;
    k_count_prefix<<<1, 1024, 0, L.stream>>>(counts, frames, prefix);
    k_pack_clouds<<<dim3(8, static_cast<unsigned>(frames)), 256, 0, L.stream>>>(slabs, counts, prefix, P, packed);
    *L.n_launches += 2;
}
}  // namespace gs
