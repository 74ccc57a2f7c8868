// pb_march_general.cu -- the general-load (planar-format leaves, partial last groups, planar / rgba8 / RGBA-f32 sinks) variants of k_fused_march (pb_march_impl.cuh), instantiated in a translation unit of
// their own so that they compile beside the fast variants.
#include "pb_march_impl.cuh"

namespace pb {

cudaError_t launch_fused_march_planar(cudaStream_t s, const FusedDesc &d, int num_sms, size_t smem, int plain, bool single) {
	auto launch = [&](void (*kernel)(const FusedDesc)) -> cudaError_t { return march_launch(kernel, s, d, num_sms, smem, kGeneralWarps); };
	const bool extras = d.feat != 0;   // Lanczos-in-launch / Yadif leaves / RGBA-f32 sink: the instances that carry them
		if (d.feat == 2 && plain != 2) {   // RGBA-f32 / Yadif leaves only (the frames of a de-interlacing channel, routed layers)
			if (plain) return single ? launch(k_fused_march<1, true, true, 1, true, false, false, 2>) : launch(k_fused_march<1, true, false, 1, true, false, false, 2>);
			return launch(k_fused_march<1, true, false, 0, true, false, false, 2>);
		}
		if (extras) {
			if (plain == 2) return single ? launch(k_fused_march<1, true, true, 2, true, false, false, 7>) : launch(k_fused_march<1, true, false, 2, true, false, false, 7>);
			if (plain) return single ? launch(k_fused_march<1, true, true, 1, true, false, false, 7>) : launch(k_fused_march<1, true, false, 1, true, false, false, 7>);
			return launch(k_fused_march<1, true, false, 0, true, false, false, 7>);
		}
		if (plain == 2) return single ? launch(k_fused_march<1, true, true, 2, true>) : launch(k_fused_march<1, true, false, 2, true>);
		if (plain) return single ? launch(k_fused_march<1, true, true, 1, true>) : launch(k_fused_march<1, true, false, 1, true>);
		return launch(k_fused_march<1, true, false, 0, true>);
}

}  // namespace pb
