// pb_march_bigrows.cu -- the big-row (64 source groups per strip row; rgba8 / bgra8 / RGBA-f32 / Yadif leaves) variants of k_fused_march (pb_march_impl.cuh), instantiated in a translation unit of
// their own so that they compile beside the fast variants.
#include "pb_march_impl.cuh"

namespace pb {

cudaError_t launch_fused_march_bigrows(cudaStream_t s, const FusedDesc &d, int num_sms, size_t smem, int plain, bool single) {
	auto launch = [&](void (*kernel)(const FusedDesc)) -> cudaError_t { return march_launch(kernel, s, d, num_sms, smem, kGeneralWarps); };
	const bool extras = d.feat != 0;   // Lanczos-in-launch / Yadif leaves / RGBA-f32 sink: the instances that carry them
		if (d.feat == 2 && plain != 2) {   // RGBA-f32 / Yadif leaves only (the frames of a de-interlacing channel, routed layers)
			if (plain) return single ? launch(k_fused_march<1, true, true, 1, true, true, false, 2>) : launch(k_fused_march<1, true, false, 1, true, true, false, 2>);
			return launch(k_fused_march<1, true, false, 0, true, true, false, 2>);
		}
		if (extras) {
			if (plain == 2) return single ? launch(k_fused_march<1, true, true, 2, true, true, false, 7>) : launch(k_fused_march<1, true, false, 2, true, true, false, 7>);
			if (plain) return single ? launch(k_fused_march<1, true, true, 1, true, true, false, 7>) : launch(k_fused_march<1, true, false, 1, true, true, false, 7>);
			return launch(k_fused_march<1, true, false, 0, true, true, false, 7>);
		}
		if (plain == 2) return single ? launch(k_fused_march<1, true, true, 2, true, true>) : launch(k_fused_march<1, true, false, 2, true, true>);
		if (plain) return single ? launch(k_fused_march<1, true, true, 1, true, true>) : launch(k_fused_march<1, true, false, 1, true, true>);
		return launch(k_fused_march<1, true, false, 0, true, true>);
}

}  // namespace pb
