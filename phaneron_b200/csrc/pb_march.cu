// pb_march.cu -- the fast fused kernel: warp-autonomous strip marching.  (Device code: pb_march_impl.cuh; this unit
// instantiates the fast variants, the dedicated single-layer / direct kernels and the Lanczos first pass, and dispatches.)
//
//   N layers x (v210 unpack -> YCbCr->R'G'B' -> gamma LUT -> gamut 3x3 -> bilinear Transform
//   -> dissolve | wipe) -> combine (premultiplied over) -> linear->gamma LUT -> RGB->YCbCr
//   -> 10-bit RTE -> v210 pack, one launch, every packed source byte read once from HBM.
// Reference stages replaced: v210.ts:25-195, transform.ts:36-59, transition.ts:60-73,
// combine.ts:24-68 and the RGBA-f32 frames between them.
//
// Shape of the kernel:
//   * one persistent CTA per SM, kMarchWarps warps; the gamma tables live in shared memory in the
//     lossless one-byte-per-entry form of pb_lut.cuh (128 KiB for a read + a write table);
//   * a work item is one output line of one strip (15 or 16 v210 groups = 90 / 96 px, 3 pixels per
//     lane); items are dealt round-robin to all warps of the grid, so a warp never waits on another
//     warp: no __syncthreads after the table load, only __syncwarp.  Narrow strips keep the per-lane
//     pixel state small (registers, not occupancy, were the limit of wider strips: profiles/);
//   * per leaf the warp converts the strip's source footprint ONCE (lane = v210 group: one 128-bit
//     load, 6 texels) into its private planar row buffer.  At scale >= ~1 a row needs <= 16 groups,
//     so lanes 0-15 convert row j0 and lanes 16-31 row j0+1 in ONE full-width pass and the 4-tap
//     bilinear chain runs at once; smaller scales (<= 32 groups per row) take one pass per row and
//     continue the oracle's canonical chain (w00*t00 -> +w10*t10 -> +w01*t01 -> +w11*t11) across them;
//   * every lane then takes its taps for 3 output pixels (lane = pixel, stride-1 conflict-free
//     LDS) with the exact {i0, a} / {j0, b} tables the host derived from the reference's float formula;
//   * the 6-pixel / 4-word v210 regroup goes through the same buffer: 32 lanes x 3 rounds of
//     codes in, lane = group out, one coalesced 16-byte store per lane.
// Instruction economy (the kernel is issue / FMA-pipe / SFU bound, not HBM bound -- DESIGN.md 4.3):
//   * the two pixels of a 4:2:2 chroma pair are converted together with packed fp32x2
//     instructions (fma.rn.f32x2: same lane rate as FFMA, half the issue slots);
//   * 10-bit fields become floats with one LOP3 (mask | 2^23 exponent); a chroma field at bit 10
//     needs no shift: it is read as 1024*c and meets a coefficient pre-divided by 1024;
//   * the 2^23 bias of the luma floats is folded into the first FMA of the matrix row (ReadK::oY);
//   * the table index, the table address and the exact float index all come from ONE add of a
//     per-table magic constant 2^23 + (shared-memory address of the table);
//   * the toe / power select of the transfer function is arithmetic (saturating FMAs), keeping
//     the half-rate ALU pipe for the unpack masks and the final integer add.
// Every float operation is an explicit IEEE round-to-nearest op or an SFU approximation that the
// table fit has already absorbed, so results are bit-identical to the generic kernel
// (pb_fused.cu) and to the oracle.
#include "pb_march_impl.cuh"

namespace pb {

cudaError_t launch_lanczos_hpass(cudaStream_t s, const HPassDesc &h, int num_sms) {
	static std::mutex mu;
	static std::set<int> configured;
	int dev = 0;
	cudaGetDevice(&dev);
	{
		std::lock_guard<std::mutex> lk(mu);
		if (!configured.count(dev)) {
			cudaError_t e = cudaFuncSetAttribute(k_lanczos_hpass<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
			if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lanczos_hpass<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
			if (e != cudaSuccess) return e;
			configured.insert(dev);
		}
	}
	const int total = (h.j_hi - h.j_lo) * (h.s1 - h.s0 + 1);
	if (total <= 0) return cudaSuccess;
	const int grid = max(1, min(num_sms, (total + kMarchWarps - 1) / kMarchWarps));
	const size_t smem = 65536 + (size_t)kMarchWarps * kRowFloats * sizeof(float);
	if (h.lut.lp.affine == 2) k_lanczos_hpass<2><<<grid, kMarchThreads, smem, s>>>(h);
	else k_lanczos_hpass<0><<<grid, kMarchThreads, smem, s>>>(h);
	return cudaGetLastError();
}

cudaError_t launch_lut_fit(cudaStream_t s, const float *table, const LutParams *cands_dev, int n_cands, uint8_t *d8_out, void *results_dev) {
	lut_fit_kernel<<<dim3(65536 / 256, n_cands), 256, 0, s>>>(table, cands_dev, d8_out, reinterpret_cast<LutFitResult *>(results_dev));
	return cudaGetLastError();
}

size_t march_smem_bytes(const FusedDesc &d) {
	if (d.bg_single) return (size_t)d.n_luts * 65536 + (size_t)kMarchWarps * kSingleRowFloats * sizeof(float);
#ifdef PB_EXP_ROW_PREFETCH
	const size_t pf_tiles = (d.n_luts > 0 && !d.any_planar && !d.big_rows) ? (size_t)kMarchWarps * kPfBytes : 0;   // TMA row prefetch of the fast variants
#else
	const size_t pf_tiles = 0;
#endif
	const size_t warps = (d.n_luts > 0 && (d.any_planar || d.big_rows)) ? kGeneralWarps : kMarchWarps;   // (as launch_fused_march dispatches)
	return (size_t)d.n_luts * 65536 + warps * (d.big_rows ? 2 : 1) * kRowFloats * sizeof(float) + (size_t)d.n_t256 * 1024 + pf_tiles;
}

cudaError_t launch_fused_march(cudaStream_t s, const FusedDesc &d, int num_sms) {
	const size_t smem = march_smem_bytes(d);
	auto launch = [&](void (*kernel)(const FusedDesc)) -> cudaError_t { return march_launch(kernel, s, d, num_sms, smem); };
	if (d.single_lines > 0 && !d.bg_single) {   // one v210 layer through an axis-aligned Transform, vertical scale >= 1 (prepare_march checks)
		static std::mutex mu;
		static std::set<int> configured;
		int dev = 0;
		cudaGetDevice(&dev);
		{
			std::lock_guard<std::mutex> lk(mu);
			if (!configured.count(dev)) {
				cudaError_t e = cudaSuccess;
				for (auto *k : {k_march_single<false, 0>, k_march_single<true, 0>, k_march_single<false, 2>, k_march_single<true, 2>})
					if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
				if (e != cudaSuccess) return e;
				configured.insert(dev);
			}
		}
		const int n_strips = (d.out_w / 6 + d.single_strip_groups - 1) / d.single_strip_groups;
		const int total = n_strips * ((d.out_h + d.single_lines - 1) / d.single_lines);
		const int grid = max(1, min(num_sms, (total + kMarchWarps - 1) / kMarchWarps));
		const size_t smem_single = (size_t)d.n_luts * 65536 + (size_t)kMarchWarps * kSingleRowFloats * sizeof(float);
		const bool poly = d.luts[d.rc[d.layers[0].a.rc].lut_slot].lp.affine == 2;   // the layer's read table in the MUFU-free model
		const bool plain_fmt = d.layers[0].a.kind == LEAF_V210 && d.sink == SINK_V210;
		return launch_pdl(plain_fmt ? (poly ? k_march_single<false, 2> : k_march_single<false, 0>) : (poly ? k_march_single<true, 2> : k_march_single<true, 0>), grid,
		                  kMarchThreads, smem_single, s, d);
	}
	if (d.direct_mode) {   // one v210 source 1:1 into a v210 output (prepare_march checks the conditions)
		static std::mutex mu;
		static std::set<int> configured;
		int dev = 0;
		cudaGetDevice(&dev);
		{
			std::lock_guard<std::mutex> lk(mu);
			if (!configured.count(dev)) {
				cudaError_t e = cudaSuccess;
				for (auto *k : {k_march_direct<0, false>, k_march_direct<2, false>, k_march_direct<0, true>, k_march_direct<2, true>, k_march_direct<0, false, true>})
					if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
				if (e != cudaSuccess) return e;
				configured.insert(dev);
			}
		}
		const int n_lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
		const int total = n_lines * ((d.out_w / 6 + 31) / 32);
		const int grid = max(1, min(num_sms, (total + kDirectWarps - 1) / kDirectWarps));
		const size_t smem_direct = (size_t)d.n_luts * 65536 + (size_t)kDirectWarps * kRowFloats * sizeof(float);
		if (d.direct_mode == 2) {   // an RGBA-f32 frame packed to v210
			return launch_pdl(k_march_direct<0, false, true>, grid, kDirectWarps * 32, smem_direct, s, d);
		}
		const bool poly = d.luts[d.rc[0].lut_slot].lp.affine == 2;
		if (d.sink == SINK_RGBA_F32) return launch_pdl(poly ? k_march_direct<2, true> : k_march_direct<0, true>, grid, kDirectWarps * 32, smem_direct, s, d);
		return launch_pdl(poly ? k_march_direct<2, false> : k_march_direct<0, false>, grid, kDirectWarps * 32, smem_direct, s, d);
	}
	const bool single = d.n_rc == 1;
	if (d.n_luts > 0) {
		// plain: the write table is the affine MUFU model and the read tables are all the non-affine MUFU model (1) or all the
		// MUFU-free polynomial model (2): the selects inside lut2 become compile-time constants
		int plain = d.wlp.affine == 1 ? -1 : 0;
		for (int i = 0; i < d.n_rc && plain; ++i) {
			if (d.rc[i].lut_slot < 0) continue;   // (< 0: constants of rgba8 leaves only)
			const int m = d.luts[d.rc[i].lut_slot].lp.affine == 0 ? 1 : d.luts[d.rc[i].lut_slot].lp.affine == 2 ? 2 : 0;
			plain = (plain < 0 || plain == m) ? m : 0;
		}
		if (plain < 0) plain = 1;   // no YCbCr read table at all
		if (d.big_rows) return launch_fused_march_bigrows(s, d, num_sms, smem, plain, single);   // pb_march_bigrows.cu
		if (d.any_planar) return launch_fused_march_planar(s, d, num_sms, smem, plain, single);   // pb_march_general.cu
		if (d.bg_single) {   // (prepare_march: plain tables, sparse matrices, v210 leaves with whole groups)
			if (plain == 2) return single ? launch(k_fused_march<1, true, true, 2, false, false, true>) : launch(k_fused_march<1, true, false, 2, false, false, true>);
			return single ? launch(k_fused_march<1, true, true, 1, false, false, true>) : launch(k_fused_march<1, true, false, 1, false, false, true>);
		}
		if (plain == 2 && d.sparse_cm) return single ? launch(k_fused_march<1, true, true, 2>) : launch(k_fused_march<1, true, false, 2>);
		if (plain && d.sparse_cm) return single ? launch(k_fused_march<1, true, true, 1>) : launch(k_fused_march<1, true, false, 1>);
		if (d.sparse_cm) return single ? launch(k_fused_march<1, true, true>) : launch(k_fused_march<1, true, false>);
		return single ? launch(k_fused_march<1, false, true>) : launch(k_fused_march<1, false, false>);
	}
	if (d.sparse_cm) return single ? launch(k_fused_march<0, true, true>) : launch(k_fused_march<0, true, false>);
	return single ? launch(k_fused_march<0, false, true>) : launch(k_fused_march<0, false, false>);
}

}  // namespace pb
