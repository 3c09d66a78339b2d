// Hand-written sm_100a kernels of the frame-preparation hot path.
//
//   K1  fused_prepare_kernel   unMapImage (PhotometricUndistorter.cpp:193-211) composed with
//                              undistort<T> (FOVUndistorter.cpp:341-367) as DatasetReader::getImage
//                              does (BenchmarkDatasetReader.h:210-241), plus pyramid levels 1..4 in
//                              the epilogue.  Batched over frames.
//   K1a unmap_kernel           unMapImage alone (photo-only mode, BenchmarkDatasetReader.h:216)
//   K1b undistort_f32_kernel   undistort<float> alone (API parity, FOVUndistorter.cpp:369)
//   K2  pyr_down_kernel        one stand-alone pyramid level
//   K3  estep_kernel           responseCalib E-step (main_responseCalib.cpp:324-338)
//
// Arithmetic policy: every floating-point operation of the reference's loops is issued with an
// explicit round-to-nearest intrinsic (__fmul_rn/__fadd_rn/__fsub_rn, __dmul_rn/__dadd_rn/__ddiv_rn)
// in the reference's evaluation order, so nvcc cannot contract them into FMAs and the results are
// bit-identical to the reference's non-FMA x86-64 build (only NaN payloads may differ).
//
// All of this is HBM-bound byte/float streaming; there is no GEMM-shaped work, so no tcgen05.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdc_kernels.cuh"

namespace mdc {

// =====================================================================================
// small PTX helpers: mbarrier + TMA (cp.async.bulk.tensor)
// =====================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled TMA load: box (bw, bh, 1) of the u8 frame stack at (x, y, frame) -> shared memory.
__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int frame) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(frame), "r"(smem_u32(bar))
        : "memory");
}

// =====================================================================================
// K1: fused photometric un-map + FOV rectification (+ pyramid epilogue)
// =====================================================================================
//
// Work decomposition.  The output image is cut into 32x32 tiles; a tile's bilinear taps fall into
// a compact bounding box of the input image (host-precomputed, TileDesc).  The unit of work is
// (tile, frame).  Units are ordered tile-major and split into gridDim.x contiguous, cost-balanced
// ranges, one per persistent CTA, so a CTA works through a few tiles and, for each, loops over a
// long run of frames.  Everything that depends only on the calibration — the remap entry of each of
// the thread's 4 pixels (turned into 4 bilinear weights and a box-local offset) and the 4 vignette
// reciprocals under the taps — is loaded ONCE per tile into registers and reused for every frame,
// which removes the 15.7 MB/frame of table traffic a naive fused kernel would add to the
// 6.5 MB/frame of unavoidable image traffic (SURVEY.md §7 "hard parts").
//
// Per frame: the u8 input box is brought on chip (TMA tensor load into a 2-stage mbarrier ring, or
// register-prefetched LDG for image widths TMA cannot describe), pushed once through the
// response LUT (lane-replicated in shared memory, so the 256-entry gather is bank-conflict free)
// into a float tile, and then each thread gathers 4 taps per pixel from that tile, applies the
// vignette reciprocals and the bilinear blend in the reference's exact operation order, and
// writes its 2x2 block.  The 2x2 ownership makes pyramid level 1 thread-local, level 2 a 4-lane
// shuffle, and levels 3-4 a 64-float shared-memory hand-off to warp 0.

struct TileRegs {
    float w[4][4];    // per pixel: w0 (x,y) w1 (x+1,y) w2 (x,y+1) w3 (x+1,y+1)
    float vi[4][4];   // vignette reciprocals under the same taps
    int off[4];       // box-local (staged) or image-global (direct) offset of tap 0; <0 = black pixel
};

__device__ __forceinline__ float lut_value(const FusedParams& p, int v) {
    float r = p.lut_gamma ? __ldg(p.ginv + v) : static_cast<float>(v);
    if (p.kill && v == 255) r = __int_as_float(0x7fc00000);  // NAN
    return r;
}

// locate the (tile, frame) at work position `pos` (in cost units) of the tile-major unit order
__device__ void locate_unit(const FusedParams& p, unsigned long long pos, int& tile, int& frame) {
    const unsigned long long nf = static_cast<unsigned long long>(p.n_frames);
    const unsigned long long total = static_cast<unsigned long long>(p.tile_cost_prefix[p.n_tiles]) * nf;
    if (pos >= total) { tile = p.n_tiles; frame = 0; return; }
    int lo = 0, hi = p.n_tiles;  // invariant: prefix[lo]*nf <= pos < prefix[hi]*nf
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (static_cast<unsigned long long>(p.tile_cost_prefix[mid]) * nf <= pos) lo = mid; else hi = mid;
    }
    const unsigned long long base = static_cast<unsigned long long>(p.tile_cost_prefix[lo]) * nf;
    const unsigned long long wt = p.tile_cost_prefix[lo + 1] - p.tile_cost_prefix[lo];
    tile = lo;
    frame = static_cast<int>((pos - base) / wt);
}

template <bool kTma>
__global__ void __launch_bounds__(kThreads, 2) fused_prepare_kernel(const __grid_constant__ FusedParams p, const __grid_constant__ TmaMaps maps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // ---- shared-memory carve-up (see fused_smem_bytes)
    float* lut = reinterpret_cast<float*>(smem_raw);                    // [256][32] lane-replicated response LUT
    float* ftile = lut + 256 * 32;                                      // [box_px_max] LUT-mapped input box
    float* s_l2 = ftile + p.box_px_max;                                 // [64] pyramid level-2 hand-off
    int* s_sched = reinterpret_cast<int*>(s_l2 + 64);                   // [4] tile_b, frame_b, tile_e, frame_e
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_sched + 4);         // [kStages] mbarriers (TMA only)
    const uint32_t stage_bytes = (static_cast<uint32_t>(p.box_px_max) + 127u) & ~127u;
    uint8_t* u8stage = smem_raw + ((256u * 32u * 4u + static_cast<uint32_t>(p.box_px_max) * 4u + 64u * 4u + 16u + 8u * kStages + 127u) & ~127u);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bx = tid & 15, by = tid >> 4;   // position of this thread's 2x2 block inside the tile

    // ---- one-time per CTA: LUT, barriers, work range
    for (int i = tid; i < 256 * 32; i += kThreads) lut[i] = lut_value(p, i >> 5);
    if (tid == 0) {
        const unsigned long long nf = static_cast<unsigned long long>(p.n_frames);
        const unsigned long long total = static_cast<unsigned long long>(p.tile_cost_prefix[p.n_tiles]) * nf;
        int t, f;
        locate_unit(p, total * blockIdx.x / gridDim.x, t, f);
        s_sched[0] = t; s_sched[1] = f;
        locate_unit(p, total * (blockIdx.x + 1ull) / gridDim.x, t, f);
        s_sched[2] = t; s_sched[3] = f;
        if (kTma) {
            for (int s = 0; s < kStages; ++s) mbar_init(&s_bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const int tile_b = s_sched[0], frame_b = s_sched[1], tile_e = s_sched[2], frame_e = s_sched[3];
    const size_t n_in = static_cast<size_t>(p.in_w) * p.in_h;
    const size_t n_out0 = static_cast<size_t>(p.lw[0]) * p.lh[0];
    uint32_t it = 0;  // frames consumed so far by this CTA (TMA stage/parity bookkeeping)

    for (int tile = tile_b; tile <= tile_e && tile < p.n_tiles; ++tile) {
        const int f_begin = (tile == tile_b) ? frame_b : 0;
        const int f_end = (tile == tile_e) ? frame_e : p.n_frames;
        if (f_begin >= f_end) continue;

        // ------------------------------------------------------------ per-tile prologue
        const TileDesc td = p.tiles[tile];
        const int mode = td.mode_map & 0xff;
        const int bw = td.bw_bh & 0xffff, bh = td.bw_bh >> 16;
        const int tx0 = (tile % p.tiles_x) * kTile, ty0 = (tile / p.tiles_x) * kTile;
        TileRegs r;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ox = tx0 + 2 * bx + (q & 1), oy = ty0 + 2 * by + (q >> 1);
            float sx = -1.0f, sy = -1.0f;
            if (ox < p.out_w && oy < p.out_h) {
                const size_t o = static_cast<size_t>(oy) * p.out_w + ox;
                sx = __ldg(p.remap_x + o);
                sy = __ldg(p.remap_y + o);
            }
            r.off[q] = -1;
#pragma unroll
            for (int k = 0; k < 4; ++k) { r.w[q][k] = 0.0f; r.vi[q][k] = 1.0f; }
            if (!(sx < 0)) {  // the reference tests only remapX (FOVUndistorter.cpp:347)
                const int xi = static_cast<int>(sx), yi = static_cast<int>(sy);   // truncation, :352-353
                const float fx = __fsub_rn(sx, static_cast<float>(xi));
                const float fy = __fsub_rn(sy, static_cast<float>(yi));
                const float fxy = __fmul_rn(fx, fy);
                r.w[q][3] = fxy;
                r.w[q][2] = __fsub_rn(fy, fxy);
                r.w[q][1] = __fsub_rn(fx, fxy);
                r.w[q][0] = __fadd_rn(__fsub_rn(__fsub_rn(1.0f, fx), fy), fxy);
                const int g = yi * p.in_w + xi;
                r.off[q] = (mode == TILE_STAGED) ? (yi - td.y0) * bw + (xi - td.x0) : g;
                if (p.use_vig) {
                    r.vi[q][0] = __ldg(p.vinv + g);
                    r.vi[q][1] = __ldg(p.vinv + g + 1);
                    r.vi[q][2] = __ldg(p.vinv + g + p.in_w);
                    r.vi[q][3] = __ldg(p.vinv + g + p.in_w + 1);
                }
            }
        }

        // loader geometry for the staged modes
        const int n_words = (bw * bh) >> 2;                 // u32 words in the box
        // LDG loader: threads tiled (rows x words-per-row) with a power-of-two row length
        int lg = 2;                                          // log2 of padded words per row (>= 4 words)
        while ((4 << lg) < bw) ++lg;
        const int ld_col = tid & ((1 << lg) - 1), ld_row = tid >> lg, ld_rstep = kThreads >> lg;
        const int ld_pass = (mode == TILE_STAGED && !kTma) ? (bh + ld_rstep - 1) / ld_rstep : 0;
        const bool ld_col_ok = (ld_col * 4) < bw;
        const bool ld_fast = ((p.in_w & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.frames) & 3) == 0);
        uint32_t pre[kMaxBoxWordsPerThread];

        auto ldg_box = [&](int frame) {   // issue the global loads of `frame`'s box into registers
            const uint8_t* src = p.frames + static_cast<size_t>(frame) * n_in;
#pragma unroll
            for (int k = 0; k < kMaxBoxWordsPerThread; ++k) {
                uint32_t v = 0;
                if (k < ld_pass) {
                    const int row = ld_row + k * ld_rstep, gy = td.y0 + row, gx = td.x0 + ld_col * 4;
                    if (ld_col_ok && row < bh && gy < p.in_h && gx < p.in_w) {
                        const uint8_t* a = src + static_cast<size_t>(gy) * p.in_w + gx;
                        if (ld_fast) v = __ldg(reinterpret_cast<const uint32_t*>(a));
                        else {
                            v = __ldg(a);
                            if (gx + 1 < p.in_w) v |= static_cast<uint32_t>(__ldg(a + 1)) << 8;
                            if (gx + 2 < p.in_w) v |= static_cast<uint32_t>(__ldg(a + 2)) << 16;
                            if (gx + 3 < p.in_w) v |= static_cast<uint32_t>(__ldg(a + 3)) << 24;
                        }
                    }
                }
                pre[k] = v;
            }
        };
        auto lut4 = [&](uint32_t v) {
            float4 f;
            f.x = lut[((v & 0xffu) << 5) + lane];
            f.y = lut[(((v >> 8) & 0xffu) << 5) + lane];
            f.z = lut[(((v >> 16) & 0xffu) << 5) + lane];
            f.w = lut[((v >> 24) << 5) + lane];
            return f;
        };

        const CUtensorMap* tmap = &maps.m[kTma ? ((td.mode_map >> 8) & 0xff) : 0];
        const uint32_t tma_bytes = static_cast<uint32_t>(bw) * static_cast<uint32_t>(td.mode_map >> 16);
        if (mode == TILE_STAGED) {
            if (kTma) {
                if (tid == 0) {
#pragma unroll
                    for (int s = 0; s < kStages; ++s)
                        if (f_begin + s < f_end) {
                            const uint32_t st = (it + s) % kStages;
                            mbar_expect_tx(&s_bar[st], tma_bytes);
                            tma_load_box(u8stage + st * stage_bytes, tmap, &s_bar[st], td.x0, td.y0, f_begin + s);
                        }
                }
            } else {
                ldg_box(f_begin);
            }
        }

        // ------------------------------------------------------------ frame loop
        for (int f = f_begin; f < f_end; ++f) {
            const uint8_t* frame = p.frames + static_cast<size_t>(f) * n_in;
            if (mode == TILE_STAGED) {
                // bring the box through the response LUT into the float tile
                if (kTma) {
                    const uint32_t st = it % kStages;
                    mbar_wait(&s_bar[st], (it / kStages) & 1u);
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(u8stage + st * stage_bytes);
                    float4* dst = reinterpret_cast<float4*>(ftile);
                    for (int g = tid; g < n_words; g += kThreads) dst[g] = lut4(src[g]);
                } else {
#pragma unroll
                    for (int k = 0; k < kMaxBoxWordsPerThread; ++k) {
                        const int row = ld_row + k * ld_rstep;
                        if (k < ld_pass && ld_col_ok && row < bh)
                            *reinterpret_cast<float4*>(ftile + row * bw + ld_col * 4) = lut4(pre[k]);
                    }
                }
            }
            __syncthreads();   // (A) float tile complete; u8 stage / prefetch registers free
            if (mode == TILE_STAGED) {
                if (kTma) {
                    if (tid == 0 && f + kStages < f_end) {
                        const uint32_t st = it % kStages;
                        mbar_expect_tx(&s_bar[st], tma_bytes);
                        tma_load_box(u8stage + st * stage_bytes, tmap, &s_bar[st], td.x0, td.y0, f + kStages);
                    }
                } else if (f + 1 < f_end) {
                    ldg_box(f + 1);   // in flight during the gather below
                }
            }

            // ---- gather + blend (reference order: FOVUndistorter.cpp:362-365)
            float px[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v = 0.0f;
                if (r.off[q] >= 0) {
                    float g0, g1, g2, g3;
                    if (mode == TILE_STAGED) {
                        const float* s = ftile + r.off[q];
                        g0 = s[0]; g1 = s[1]; g2 = s[bw]; g3 = s[bw + 1];
                    } else {
                        const uint8_t* s = frame + r.off[q];
                        g0 = lut[(static_cast<int>(__ldg(s)) << 5) + lane];
                        g1 = lut[(static_cast<int>(__ldg(s + 1)) << 5) + lane];
                        g2 = lut[(static_cast<int>(__ldg(s + p.in_w)) << 5) + lane];
                        g3 = lut[(static_cast<int>(__ldg(s + p.in_w + 1)) << 5) + lane];
                    }
                    if (p.use_vig) {   // unMapImage: GInv[I] * vignetteMapInv (PhotometricUndistorter.cpp:205)
                        g0 = __fmul_rn(g0, r.vi[q][0]); g1 = __fmul_rn(g1, r.vi[q][1]);
                        g2 = __fmul_rn(g2, r.vi[q][2]); g3 = __fmul_rn(g3, r.vi[q][3]);
                    }
                    v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.w[q][3], g3), __fmul_rn(r.w[q][2], g2)),
                                            __fmul_rn(r.w[q][1], g1)),
                                  __fmul_rn(r.w[q][0], g0));
                }
                px[q] = v;
            }

            // ---- level 0 store (2x2 block)
            {
                const int ox = tx0 + 2 * bx, oy = ty0 + 2 * by;
                float* o = p.out[0] + static_cast<size_t>(f) * n_out0 + static_cast<size_t>(oy) * p.out_w + ox;
                if (p.vec2_ok && ox + 1 < p.out_w) {
                    if (oy < p.out_h) *reinterpret_cast<float2*>(o) = make_float2(px[0], px[1]);
                    if (oy + 1 < p.out_h) *reinterpret_cast<float2*>(o + p.out_w) = make_float2(px[2], px[3]);
                } else {
                    if (oy < p.out_h) {
                        if (ox < p.out_w) o[0] = px[0];
                        if (ox + 1 < p.out_w) o[1] = px[1];
                    }
                    if (oy + 1 < p.out_h) {
                        if (ox < p.out_w) o[p.out_w] = px[2];
                        if (ox + 1 < p.out_w) o[p.out_w + 1] = px[3];
                    }
                }
            }

            // ---- pyramid epilogue: dst = 0.25f*(((a+b)+c)+d), a=(2x,2y) b=(2x+1,2y) c=(2x,2y+1) d=(2x+1,2y+1)
            if (p.levels > 1) {
                const float l1 = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(px[0], px[1]), px[2]), px[3]));
                {
                    const int X = (tx0 >> 1) + bx, Y = (ty0 >> 1) + by;
                    if (X < p.lw[1] && Y < p.lh[1])
                        p.out[1][static_cast<size_t>(f) * p.lw[1] * p.lh[1] + static_cast<size_t>(Y) * p.lw[1] + X] = l1;
                }
                if (p.levels > 2) {
                    // a 2x2 group of level-1 pixels lives in lanes (l, l+1, l+16, l+17) of one warp
                    const float b = __shfl_down_sync(0xffffffffu, l1, 1);
                    const float c = __shfl_down_sync(0xffffffffu, l1, 16);
                    const float d = __shfl_down_sync(0xffffffffu, l1, 17);
                    const float l2 = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(l1, b), c), d));
                    if (lane < 16 && (lane & 1) == 0) {
                        const int X = (tx0 >> 2) + (bx >> 1), Y = (ty0 >> 2) + warp;
                        if (X < p.lw[2] && Y < p.lh[2])
                            p.out[2][static_cast<size_t>(f) * p.lw[2] * p.lh[2] + static_cast<size_t>(Y) * p.lw[2] + X] = l2;
                        s_l2[warp * 8 + (bx >> 1)] = l2;
                    }
                }
            }
            __syncthreads();   // (B) float tile free for the next frame; level-2 hand-off visible
            if (kTma && mode == TILE_STAGED) ++it;   // one mbarrier phase consumed
            if (p.levels > 3 && warp == 0) {
                float l3 = 0.0f;
                if (lane < 16) {
                    const int X = lane & 3, Y = lane >> 2;
                    const float* s = s_l2 + (2 * Y) * 8 + 2 * X;
                    l3 = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(s[0], s[1]), s[8]), s[9]));
                    const int GX = (tx0 >> 3) + X, GY = (ty0 >> 3) + Y;
                    if (GX < p.lw[3] && GY < p.lh[3])
                        p.out[3][static_cast<size_t>(f) * p.lw[3] * p.lh[3] + static_cast<size_t>(GY) * p.lw[3] + GX] = l3;
                }
                if (p.levels > 4) {
                    const float b = __shfl_down_sync(0xffffffffu, l3, 1);
                    const float c = __shfl_down_sync(0xffffffffu, l3, 4);
                    const float d = __shfl_down_sync(0xffffffffu, l3, 5);
                    const float l4 = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(l3, b), c), d));
                    if (lane < 16 && (lane & 1) == 0 && (lane & 4) == 0) {
                        const int GX = (tx0 >> 4) + ((lane & 3) >> 1), GY = (ty0 >> 4) + (lane >> 3);
                        if (GX < p.lw[4] && GY < p.lh[4])
                            p.out[4][static_cast<size_t>(f) * p.lw[4] * p.lh[4] + static_cast<size_t>(GY) * p.lw[4] + GX] = l4;
                    }
                }
            }
        }
    }
}

// smem layout (bytes): lut 32768 | ftile 4*box | s_l2 256 | sched 16 | bars 8*kStages | pad to 128 | stages
size_t fused_smem_bytes(int box_px_max, bool tma) {
    size_t b = 256u * 32u * 4u + static_cast<size_t>(box_px_max) * 4u + 64u * 4u + 16u + 8u * kStages;
    b = (b + 127u) & ~static_cast<size_t>(127u);
    if (tma) b += static_cast<size_t>(kStages) * ((static_cast<size_t>(box_px_max) + 127u) & ~static_cast<size_t>(127u));
    return b;
}

int fused_max_ctas_per_sm(int box_px_max, bool tma) {
    int n = 0;
    const int smem = static_cast<int>(fused_smem_bytes(box_px_max, tma));
    // the opt-in limit must be raised before the occupancy query, or it reports 0 for > 48 KB
    cudaError_t e = tma ? cudaFuncSetAttribute(fused_prepare_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                        : cudaFuncSetAttribute(fused_prepare_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return 0;
    e = tma ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_prepare_kernel<true>, kThreads, smem)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_prepare_kernel<false>, kThreads, smem);
    return e == cudaSuccess ? n : 0;
}

cudaError_t launch_fused(const FusedParams& p, const TmaMaps* maps, int grid, cudaStream_t stream) {
    const bool tma = maps != nullptr;
    const size_t smem = fused_smem_bytes(p.box_px_max, tma);
    cudaError_t e;
    if (tma) {
        e = cudaFuncSetAttribute(fused_prepare_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        fused_prepare_kernel<true><<<grid, kThreads, smem, stream>>>(p, *maps);
    } else {
        static const TmaMaps none = {};
        e = cudaFuncSetAttribute(fused_prepare_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        fused_prepare_kernel<false><<<grid, kThreads, smem, stream>>>(p, none);
    }
    return cudaGetLastError();
}

// =====================================================================================
// K1a: unMapImage alone (streaming; 1 B in, 4 B out, 4 B vignette per pixel)
// =====================================================================================
__global__ void __launch_bounds__(256) unmap_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, size_t n,
                                                    int n_frames, const float* __restrict__ ginv,
                                                    const float* __restrict__ vinv, unsigned kill) {
    __shared__ float lut[256];
    {
        float v = ginv ? ginv[threadIdx.x] : static_cast<float>(threadIdx.x);
        if (kill && threadIdx.x == 255) v = __int_as_float(0x7fc00000);
        lut[threadIdx.x] = v;
    }
    __syncthreads();
    const size_t total = n * static_cast<size_t>(n_frames);
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                     (!vinv || (reinterpret_cast<uintptr_t>(vinv) & 15) == 0);
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    if (vec) {
        const size_t words = total >> 2, nw = n >> 2;
        for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < words; i += stride) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(in) + i);
            float4 o = make_float4(lut[v & 0xff], lut[(v >> 8) & 0xff], lut[(v >> 16) & 0xff], lut[v >> 24]);
            if (vinv) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(vinv) + (i % nw));
                o.x = __fmul_rn(o.x, m.x); o.y = __fmul_rn(o.y, m.y); o.z = __fmul_rn(o.z, m.z); o.w = __fmul_rn(o.w, m.w);
            }
            reinterpret_cast<float4*>(out)[i] = o;
        }
    } else {
        for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
            float o = lut[in[i]];
            if (vinv) o = __fmul_rn(o, __ldg(vinv + (i % n)));
            out[i] = o;
        }
    }
}

cudaError_t launch_unmap(const uint8_t* in, float* out, size_t n, int n_frames, const float* ginv, const float* vinv,
                         unsigned kill, cudaStream_t stream) {
    const size_t total = n * static_cast<size_t>(n_frames);
    if (total == 0) return cudaSuccess;
    size_t blocks = (total / 4 + 255) / 256;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    if (blocks < 1) blocks = 1;
    unmap_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(in, out, n, n_frames, ginv, vinv, kill);
    return cudaGetLastError();
}

// =====================================================================================
// K1b: undistort<float> alone — float image in, tables from L2, one thread per output pixel
// =====================================================================================
__global__ void __launch_bounds__(256) undistort_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int in_w,
                                                            size_t n_in, size_t n_out, int n_frames,
                                                            const float* __restrict__ remap_x, const float* __restrict__ remap_y) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n_out; idx += stride) {
        const float sx = __ldg(remap_x + idx), sy = __ldg(remap_y + idx);
        if (sx < 0) {
            for (int f = 0; f < n_frames; ++f) out[static_cast<size_t>(f) * n_out + idx] = 0.0f;
            continue;
        }
        const int xi = static_cast<int>(sx), yi = static_cast<int>(sy);
        const float fx = __fsub_rn(sx, static_cast<float>(xi)), fy = __fsub_rn(sy, static_cast<float>(yi));
        const float fxy = __fmul_rn(fx, fy);
        const float w3 = fxy, w2 = __fsub_rn(fy, fxy), w1 = __fsub_rn(fx, fxy);
        const float w0 = __fadd_rn(__fsub_rn(__fsub_rn(1.0f, fx), fy), fxy);
        const size_t g = static_cast<size_t>(yi) * in_w + xi;
        for (int f = 0; f < n_frames; ++f) {
            const float* s = in + static_cast<size_t>(f) * n_in + g;
            const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w3, __ldg(s + in_w + 1)), __fmul_rn(w2, __ldg(s + in_w))),
                                                __fmul_rn(w1, __ldg(s + 1))),
                                      __fmul_rn(w0, __ldg(s)));
            out[static_cast<size_t>(f) * n_out + idx] = v;
        }
    }
}

cudaError_t launch_undistort_f32(const float* in, float* out, int in_w, int n_in, int n_out, int n_frames,
                                 const float* remap_x, const float* remap_y, cudaStream_t stream) {
    if (n_out == 0 || n_frames == 0) return cudaSuccess;
    int blocks = (n_out + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    undistort_f32_kernel<<<blocks, 256, 0, stream>>>(in, out, in_w, static_cast<size_t>(n_in), static_cast<size_t>(n_out), n_frames,
                                                     remap_x, remap_y);
    return cudaGetLastError();
}

// =====================================================================================
// K2: stand-alone pyramid level (used for levels beyond the fused epilogue and for API parity)
// =====================================================================================
__global__ void __launch_bounds__(256) pyr_down_kernel(const float* __restrict__ src, int sw, int sh, float* __restrict__ dst,
                                                       int n_frames) {
    const int dw = sw >> 1, dh = sh >> 1;
    const size_t per = static_cast<size_t>(dw) * dh, total = per * n_frames;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t f = i / per, r = i - f * per;
        const int y = static_cast<int>(r / dw), x = static_cast<int>(r - static_cast<size_t>(y) * dw);
        const float* s = src + f * static_cast<size_t>(sw) * sh + static_cast<size_t>(2 * y) * sw + 2 * x;
        dst[i] = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(__ldg(s), __ldg(s + 1)), __ldg(s + sw)), __ldg(s + sw + 1)));
    }
}

cudaError_t launch_pyr_down(const float* src, int sw, int sh, float* dst, int n_frames, cudaStream_t stream) {
    const size_t total = static_cast<size_t>(sw >> 1) * (sh >> 1) * n_frames;
    if (total == 0) return cudaSuccess;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    pyr_down_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, sw, sh, dst, n_frames);
    return cudaGetLastError();
}

// =====================================================================================
// K3: responseCalib E-step.  One thread per pixel (4 adjacent pixels per thread when the plane
// size allows 32-bit loads), sequential over the n exposures in the reference's order, fp64
// with explicit non-fused multiplies/adds  ->  bit-identical to main_responseCalib.cpp:324-338.
// =====================================================================================
template <int kPix>
__global__ void __launch_bounds__(256) estep_kernel(const uint8_t* __restrict__ data, int n, size_t npix,
                                                    const double* __restrict__ t, const double* __restrict__ G,
                                                    double* __restrict__ E) {
    __shared__ double sG[256];
    sG[threadIdx.x] = G[threadIdx.x];
    __syncthreads();
    const size_t k0 = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * kPix;
    if (k0 >= npix) return;
    double esum[kPix], enumr[kPix];
#pragma unroll
    for (int j = 0; j < kPix; ++j) { esum[j] = 0.0; enumr[j] = 0.0; }
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
        const double ti = __ldg(t + i);
        const double tt = __dmul_rn(ti, ti);
        uint32_t v;
        if (kPix == 4) v = __ldg(reinterpret_cast<const uint32_t*>(data + static_cast<size_t>(i) * npix + k0));
        else v = __ldg(data + static_cast<size_t>(i) * npix + k0);
#pragma unroll
        for (int j = 0; j < kPix; ++j) {
            const unsigned b = (v >> (8 * j)) & 0xffu;
            if (b != 255u) {
                enumr[j] = __dadd_rn(enumr[j], tt);
                esum[j] = __dadd_rn(esum[j], __dmul_rn(sG[b], ti));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kPix; ++j) {
        double e = __ddiv_rn(esum[j], enumr[j]);
        if (e < 0) e = 0;
        E[k0 + j] = e;
    }
}

cudaError_t launch_estep(const uint8_t* data, int n, int npix, const double* t, const double* G, double* E, cudaStream_t stream) {
    if (npix <= 0) return cudaSuccess;
    const bool vec = (npix % 4 == 0) && ((reinterpret_cast<uintptr_t>(data) & 3) == 0);
    if (vec) {
        const int threads = npix / 4;
        estep_kernel<4><<<(threads + 255) / 256, 256, 0, stream>>>(data, n, static_cast<size_t>(npix), t, G, E);
    } else {
        estep_kernel<1><<<(npix + 255) / 256, 256, 0, stream>>>(data, n, static_cast<size_t>(npix), t, G, E);
    }
    return cudaGetLastError();
}

}  // namespace mdc
