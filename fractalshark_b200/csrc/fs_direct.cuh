// fs_direct.cuh -- direct (non-perturbed) escape-time kernels, row a6 of SURVEY.md section 8.
//
// What: FractalSharkGpuLib/LowPrecisionKernels.cuh:290-382 (mandel_1x_double) and :680-790
// (mandel_1x_float): z <- z^2 + c with round-down FMAs, P iterations per bailout test,
// n_iterations -= P-1, output row flipped (Y -> height-Y-1).  Rounding sequence as compiled for the
// reference (sm_100a SASS): x0 = fma(X, dx, cx); per step  t = fma_rd(-y, y, x0);
// y' = fma_rd(x+x, y, y0); x' = fma_rd(x, x, t); bailout on fma(x, x, y*y) < 4.
//
// How: same persistent warp-tile queue as the perturbation kernel (fs_lav2.cuh).
#pragma once
#include "fs_types.cuh"

namespace fs {

template <class M> struct DirectOps;
template <> struct DirectOps<float> {
    FS_D static float fma_rd(float a, float b, float c) { return __fmaf_rd(a, b, c); }
    FS_D static float from_int(int x) { return (float)x; }
};
template <> struct DirectOps<double> {
    FS_D static double fma_rd(double a, double b, double c) { return __fma_rd(a, b, c); }
    FS_D static double from_int(int x) { return (double)x; }
};

template <class M, class IterT> struct DirectArgs {
    IterT *out;
    int width, height, pitch;
    int shard_count, shard_index; // 4-row tile bands are dealt round-robin to shards (multi-GPU)
    M cx, cy, dx, dy;
    IterT n_iterations;
    TileQueue queue;
    unsigned long long *step_counter;
};

template <class M, class IterT, int P>
__global__ void __launch_bounds__(256) direct_kernel(const DirectArgs<M, IterT> A) {
    const int lane = threadIdx.x & 31;
    const int tiles_x = (A.width + 7) >> 3;
    const int tiles_y = (((A.height + 3) >> 2) - A.shard_index + A.shard_count - 1) / A.shard_count;
    const unsigned int n_tiles = (unsigned int)tiles_x * (unsigned int)tiles_y;
    const IterT n_iter = A.n_iterations - (IterT)(P - 1);
    unsigned long long steps = 0;

    TileCursor cursor;
    tile_queue_begin(cursor);
    for (;;) {
        unsigned int tile;
        if (!next_tile(A.queue, cursor, n_tiles, tile)) break;
        // tiles (and with them the shard's 4-row bands) are laid out over the OUTPUT rows; the kernel's own Y runs the
        // other way (row flip, LowPrecisionKernels.cuh:309,699)
        int X, Yout;
        tile_origin(tile, tiles_x, tiles_y, A.shard_count, A.shard_index, X, Yout);
        X += lane & 7;
        Yout += lane >> 3;
        if (X >= A.width || Yout >= A.height) continue;
        const int Y = A.height - 1 - Yout;

        const M x0 = fma_(DirectOps<M>::from_int(X), A.dx, A.cx);
        const M y0 = fma_(DirectOps<M>::from_int(Y), A.dy, A.cy);
        M x = 0, y = 0;
        IterT iter = 0;
        if constexpr (P == 1) {
            // Eight steps per branch: every step's bailout test is still evaluated (folded into one predicate, NaN
            // counts as escaped), and a chunk in which any test failed is discarded and replayed step by step by the
            // loop below, so the count is the one the per-step loop produces.  7 instead of 10 instructions per
            // iteration.  Invariant at the top of a chunk: |z|^2 < 4 holds for the current (x, y).
            constexpr int K = 8;
            while (iter + (IterT)K <= n_iter && iter + (IterT)K > iter) {
                const M sx = x, sy = y;
                bool ok = true;
#pragma unroll
                for (int u = 0; u < K; u++) {
                    const M x2 = x + x;
                    const M t = DirectOps<M>::fma_rd(-y, y, x0);
                    y = DirectOps<M>::fma_rd(x2, y, y0);
                    x = DirectOps<M>::fma_rd(x, x, t);
                    ok = ok && (fma_(x, x, y * y) < M(4));
                }
                if (!ok) {
                    x = sx;
                    y = sy;
                    break;
                }
                iter += (IterT)K;
            }
        }
        while (fma_(x, x, y * y) < M(4) && iter < n_iter) {
#pragma unroll
            for (int p = 0; p < P; p++) {
                const M x2 = x + x;
                const M t = DirectOps<M>::fma_rd(-y, y, x0);
                y = DirectOps<M>::fma_rd(x2, y, y0);
                x = DirectOps<M>::fma_rd(x, x, t);
            }
            iter += P;
        }
        steps += iter;
        A.out[(size_t)Yout * A.pitch + X] = iter;
    }
    if (A.step_counter) {
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
        if (lane == 0 && steps) atomicAdd(A.step_counter, steps);
    }
}

} // namespace fs
