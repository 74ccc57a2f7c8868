// pb_desc.h -- launch descriptors shared by the runtime (host) and the kernels.
//
// A "frame expression" is what phaneron's layer graph computes for one output
// frame: leaves are packed source frames (or RGBA-f32 frames that had to be
// materialised), inner nodes are Transform / Transition / Combine, the root is a
// packed writer.  The runtime flattens an expression into a FusedDesc and hands it
// to ONE kernel launch (pb_fused.cu).
#pragma once
#include <stdint.h>
#include <vector_types.h>

namespace pb {

constexpr int kMaxLayers = 8;   // combine_N inputs (combiner.ts builds N = number of layers)
constexpr int kMaxReadConsts = 8;
constexpr int kMaxRingLeaves = 8;    // strip kernel: leaves with a shared-memory row ring
constexpr int kStripPx = 192;        // output pixels per strip (= threads per CTA = 32 v210 groups)
constexpr int kRingGroups = 66;      // source v210 groups a ring row can hold
constexpr int kRingRow = kRingGroups * 6;

enum LeafKind : int { LEAF_NONE = 0, LEAF_V210 = 1, LEAF_RGBA_F32 = 2 };
enum LayerKind : int { LAYER_DIRECT = 0, LAYER_DISSOLVE = 1, LAYER_WIPE_MASK = 2 };

// Loader constants (loadSave.ts:41-64): YCbCr->RGB 3x4, gamma->linear LUT, gamut 3x3
struct ReadConsts {
	float cm[12];
	float gamut[9];
	const float *lut;
	const uint8_t *lut_res;   // optional smem-residual form (see pb_lut.cuh); may be null
};

// Saver constants (loadSave.ts:130-150): linear->gamma LUT, RGB->YCbCr 3x4
struct WriteConsts {
	float cm[12];
	const float *lut;
};

struct Leaf {
	const void *ptr;
	int kind;          // LeafKind
	int w, h;          // source dimensions in pixels
	int pitch;         // bytes per line (v210) / unused for RGBA (w*16)
	int rc;            // index into FusedDesc::rc (v210 leaves)
	int has_xf;        // 0: sample texel (x,y) directly; 1: Transform (transform.ts:36-59)
	int xf_w, xf_h;    // dimensions of the Transform's output image
	float m[6];        // rows 0 and 1 of the 3x3 transformMatrix
	// strip kernel only: exact per-column / per-row sampling tables built on the host
	// ({i0, bits(a)} per output x, {j0, bits(b)} per output y) and the leaf's ring slot
	const int2 *col_tab;
	const int2 *row_tab;
	int ring;          // index of this leaf's row ring in shared memory, -1 if none
	int pad_;
};

struct Layer {
	int kind;          // LayerKind
	float mix;         // dissolve
	Leaf a, b, mask;
};

struct FusedDesc {
	int n_layers;
	int out_w, out_h;
	int interlace;     // packer.ts:24-28 Interlace enum: 0, 1 (top), 3 (bottom)
	int out_pitch;     // bytes per output line
	int n_rc;
	void *out;
	int n_ring, band_lines;   // strip kernel
	WriteConsts wc;
	ReadConsts rc[kMaxReadConsts];
	Layer layers[kMaxLayers];
};

}  // namespace pb
