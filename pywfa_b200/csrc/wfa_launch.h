/* wfa_launch.h -- host-callable launch wrappers of wfa_kernels.cu (internal, C++ linkage). */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace wfagpu {

struct KParams;

/* mode 0: warp-per-pair (smem ring), 1: block-per-pair (smem ring), 2: block-per-pair (HBM ring) */
cudaError_t launch_align(const KParams& P, bool two_p, bool full, int mode, bool off16, int grid, int block,
                         size_t smem, cudaStream_t st);
cudaError_t init_kernels(int smem_optin);   /* once per device, at context creation */
int align_occupancy(bool two_p, bool full, int mode, bool off16, int block, size_t smem);
size_t block_reduce_smem_bytes();

/* register-resident tier (wfa_reg.cuh): regs = packed registers per wavefront (window 64*regs) */
bool reg_tier_supported(int dx, int doe, int de, int regs);
cudaError_t launch_reg(const KParams& P, int regs, bool full, int grid, int block, size_t smem, cudaStream_t st);
int reg_occupancy(int regs, bool full, int block, size_t smem);

/* packed-halfword tier (wfa_vec.cuh): nw = warps per pair (1, 8 or 16) */
cudaError_t launch_vec(const KParams& P, bool two_p, bool full, int nw, int heur, int grid, int block, size_t smem, cudaStream_t st);
int vec_occupancy(bool two_p, bool full, int nw, int heur, int block, size_t smem);

/* several CTAs per pair (long reads): groups * ncta co-resident CTAs of 512 threads; `scratch` holds
 * grid_scratch_bytes(groups) bytes of device memory */
size_t grid_scratch_bytes(int groups);
int grid_occupancy(bool two_p, bool full, size_t smem);   /* resident CTAs per SM */
cudaError_t launch_grid(const KParams& P, bool two_p, bool full, int groups, int ncta, size_t smem, void* scratch, cudaStream_t st);

/* runs_out == nullptr: count + scan (tile_sums needs cigar_order_tiles(n)+1 entries, total in the
 * last one); otherwise gather into cig_off[n+1] (values offset by cig_base) / runs_out. */
cudaError_t launch_cigar_order(const int* nruns, const long long* runs_base, long long n,
                               long long* tile_sums, const uint32_t* runs_tmp, long long* cig_off,
                               uint32_t* runs_out, long long cig_base, cudaStream_t st);
int cigar_order_tiles(long long n);

}  // namespace wfagpu
