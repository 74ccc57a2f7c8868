// pb_desc.h -- launch descriptors shared by the runtime (host) and the kernels.
//
// A "frame expression" is what phaneron's layer graph computes for one output
// frame: leaves are packed source frames (or RGBA-f32 frames that had to be
// materialised), inner nodes are Transform / Transition / Combine, the root is a
// packed writer.  The runtime flattens an expression into a FusedDesc and hands it
// to ONE kernel launch (pb_march.cu when eligible, else pb_fused.cu).
#pragma once
#include <stdint.h>
#include <vector_types.h>

namespace pb {

constexpr int kMaxLayers = 8;   // combine_N inputs (combiner.ts builds N = number of layers)
constexpr int kMaxReadConsts = 8;
constexpr int kMaxLuts = 3;          // march kernel: distinct gamma tables resident in shared memory

// ---- march kernel geometry (pb_march.cu) ---------------------------------------------
#ifndef PB_MARCH_WARPS
#define PB_MARCH_WARPS 20
#endif
constexpr int kMarchWarps = PB_MARCH_WARPS;  // warps per CTA; one persistent CTA per SM
constexpr int kMarchThreads = kMarchWarps * 32;
// The general variants of the march kernel (planar / rgba8 / RGBA-f32 leaves, the other sinks, big rows) carry more live state:
// at 20 warps x 96 registers they spill 200-600 bytes per thread; 16 warps x 128 registers do not (2160p: FFmpeg-format 4-layer
// scene 193 -> 179 us, rgba8 overlay scene 340 -> 292 us; the fast v210 variant loses 10 % that way and keeps 20 x 96).
#ifndef PB_GENERAL_WARPS
#define PB_GENERAL_WARPS 16
#endif
constexpr int kGeneralWarps = PB_GENERAL_WARPS;
#ifndef PB_MARCH_ROUNDS
#define PB_MARCH_ROUNDS 3
#endif
constexpr int kRounds = PB_MARCH_ROUNDS;     // output pixels per lane: a strip is <= 32 * kRounds px
// Output v210 groups per strip when a leaf is bilinear-sampled: 16*kRounds/3 - 1, so that at scale 1
// the n+1 texels n pixels need span <= 16*kRounds/3 source groups for any alignment.  With kRounds = 3
// that is 16 groups per row: the two source rows of a leaf convert in ONE 32-lane pass.
constexpr int kStripGroupsXf = 16 * kRounds / 3 - 1;
constexpr int kStripGroupsDirect = 16 * kRounds / 3;   // all leaves sampled 1:1
constexpr int kRowGroups = 32 * kRounds / 3;           // source groups a warp's row buffer holds (2 rows x 16 when they fit, else 1 row)
constexpr int kRowFloats = kRowGroups * 18;            // planar R | G | B per row slot

// packed source formats a fused kernel can read directly (the reference's Reader PackImpls) + materialised RGBA-f32 frames
enum LeafKind : int {
	LEAF_NONE = 0,
	LEAF_V210 = 1,        // v210.ts
	LEAF_RGBA_F32 = 2,
	LEAF_RGBA8 = 3,       // rgba8.ts (alpha carried, through the LUT)
	LEAF_BGRA8 = 4,       // bgra8.ts
	LEAF_YUV422P10 = 5,   // yuv422p10.ts: ptr = Y, ptr_u, ptr_v
	LEAF_YUV422P8 = 6,    // yuv422p8.ts
	LEAF_YUV420P = 7,     // yuv420p.ts
	LEAF_NV12 = 8,        // nv12.ts: ptr = Y, ptr_u = interleaved chroma
	LEAF_LANCZOS_V = 10,  // second pass of a separable Lanczos Transform: ptr = the horizontally filtered rows H (RGBA-f32, h source rows x w output
	                      // columns, written by k_lanczos_hpass); value = sum_j lz_wy[j] * H(x, lz_j0[y] + j)
	LEAF_YADIF = 9        // yadifCl.ts:105-167 over three RGBA-f32 frames: ptr = cur, ptr_u = prev, ptr_v = next; Leaf::yadif = parity | tff << 1 | skipSpatial << 2.
	                      // What a kernel sees (launch_desc has run the pre-pass, yadif bit 3 set): ptr = cur, ptr_u = the interpolated lines, row j >> 1
};
enum LayerKind : int { LAYER_DIRECT = 0, LAYER_DISSOLVE = 1, LAYER_WIPE_MASK = 2 };

// Lossless shared-memory form of a 65536-entry gamma table (colourMaths.ts:130-169):
//   (for every i, by construction: the bytes are fitted on the device by the decoding code itself)
//   base(i)   = i < J ? i*kt : s * ex2(G * lg2(i*p + q)) + o      (MUFU.LG2 / MUFU.EX2)
//   table[i] == bits( base(i) ) + d8[i] - 128
// The d8 bytes are produced ON the device by the same code that decodes them (pb_lut.cuh).
// The toe / power select is arithmetic (FMA pipe, no predicate): with h = sat(i + cJ) = (i >= J),
//   base(i) = i*kt + h * (pw(i) - i*kt)          -- exactly i*kt below the knee, pw(i) to a few ulp above it
// A third, MUFU-free model (affine == 2) evaluates the power segment as a degree-7 polynomial in x = i*p + q (x in
// [-1, 1] over [J, 65535]): seven FMAs on the FMA pipe instead of MUFU.LG2 + MUFU.EX2 on the quarter-rate XU pipe.  The
// host fits the coefficients per transfer function (pb_lut_cache.cpp lut_candidates); the byte table absorbs the residue as
// for the MUFU models, and a table the polynomial misses by more than a byte falls back to the MUFU model.
constexpr int kLutPolyDeg = 7;
struct LutParams {
	float p, q, G, s, o, kt, cJ;   // cJ = 1 - J
	int affine;                    // 0: s == 1 and o == 0 (gamma -> linear direction); 1: affine; 2: polynomial power segment
	float c[kLutPolyDeg + 1];      // affine == 2: pw(x) = (((c[7] x + c[6]) x + ...) x + c[0]
};

// Loader constants (loadSave.ts:41-64): YCbCr->RGB 3x4, gamma->linear LUT, gamut 3x3
struct ReadConsts {
	float cm[12];
	float gamut[9];
	const float *lut;     // raw table in global memory (always valid)
	int lut_slot;         // march kernel: index into FusedDesc::luts, -1 if the table has no compressed form
	int t256_slot;        // march kernel, constants of rgba8 / bgra8 leaves: index of the 256-entry table (lut[c * 257]) in shared memory, else -1
};

// march kernel: ReadConsts::cm rearranged.  oY = -2^23 * mY folds the bias of the exponent-trick luma
// float into the first FMA of the chain: fma(2^23 + y, m, -2^23 * m) == RN(y * m) exactly.  A 10-bit
// chroma field at bit 10 of a word is read as the float 1024*c, so its coefficient is m / 1024 (exact:
// power of two); [c][0] multiplies plain fields, [c][1] fields scaled by 1024.
struct ReadK {
	float mY[3], oY[3], mCb[3][2], mCr[3][2];
};

// Saver constants (loadSave.ts:130-150): linear->gamma LUT, RGB->YCbCr 3x4
struct WriteConsts {
	float cm[12];
	const float *lut;
	int lut_slot;
	int pad_;
};

struct LutDesc {
	const uint8_t *d8;   // 65536 bytes in global memory, copied into shared memory by each CTA
	LutParams lp;
};

struct Leaf {
	const void *ptr;
	const void *ptr_u, *ptr_v;   // planar kinds: chroma planes
	int kind;          // LeafKind
	int w, h;          // source dimensions in pixels
	int pitch;         // bytes per line (v210) / unused for RGBA (w*16)
	int rc;            // index into FusedDesc::rc (v210 leaves)
	int yadif;         // LEAF_YADIF: parity | tff << 1 | skipSpatial << 2
	int has_xf;        // 0: sample texel (x,y) directly; 1: Transform (transform.ts:36-59)
	int xf_w, xf_h;    // dimensions of the Transform's output image
	float m[6];        // rows 0 and 1 of the 3x3 transformMatrix
	// march kernel only: exact per-column / per-row sampling tables built on the host
	// ({i0, bits(a)} per output x, {j0, bits(b)} per output y) and per-strip source footprints
	// ({flags, first source group, group count, 0} per strip; flags bit0 = strip touches the image,
	// bit1 = some tap column of the strip lies outside the image)
	const int2 *col_tab;
	const int2 *row_tab;
	const int4 *strip_tab;
	// Lanczos filter (not in the reference; definition in oracle/oracle.c): lz_tx / lz_ty taps per axis, host-built tables
	// {first tap index} per output column / line and lz_t* normalised weights each.  0 taps = the reference's bilinear sampler.
	int lz_tx, lz_ty;
	const int *lz_i0, *lz_j0;
	const float *lz_wx, *lz_wy;
	int lz_sep;            // march launch: this Lanczos leaf is evaluated separably (launch_desc runs k_lanczos_hpass first and hands the kernel a LEAF_LANCZOS_V)
	const float *lz_wxt;   // march kernel: the horizontal weights tap-major, [tap][output column]: a warp's lanes (consecutive columns) read one line
	// strips [s0, s1] and output lines [y0, y1] outside of which every tap of this leaf is a border
	// texel: the kernel skips the leaf there without touching memory
	int s0, s1, y0, y1;
	// RGBA-f32 leaves: the raw gamma table of the packed source this frame was made from by the library itself (a rotated source
	// made real by the recorder), else null.  Host only: a frame known to hold table values times a gamut matrix is finite, so
	// the launch may keep its occlusion culling (NaN * 0 != 0 is the reason frames of unknown origin switch it off).
	const float *finite_lut;
};

struct Layer {
	int kind;          // LayerKind
	float mix;         // dissolve
	Leaf a, b, mask;
};

// march kernel: the layer graph flattened by the host into a list of leaf evaluations, each followed by an action
enum MarchAct : int {
	ACT_OVER = 0,        // direct layer: acc = fma(acc, 1 - p.a, p)                          (combine.ts:49-59)
	ACT_DIS_B = 1,       // dissolve, second input first: t = p * (1 - mix)                   (transition.ts:60-65)
	ACT_DIS_A_OVER = 2,  // dissolve, first input: p = fma(p, mix, t); then over
	ACT_WIPE_M = 3,      // wipe, mask first: m = p.r                                         (transition.ts:66-73)
	ACT_WIPE_A = 4,      // wipe, first input: t = p * (1 - m)
	ACT_WIPE_B_OVER = 5  // wipe, second input: p = fma(p, m, t); then over
};
struct MarchOp {
	int layer, which;    // the leaf: (&layers[layer].a)[which]
	int act;             // MarchAct
	float mix;
};
constexpr int kMaxOps = 3 * kMaxLayers;   // 24: op masks share a word with 8 layer-opacity bits
constexpr int kMaxStrips = 128;

// what a fused launch writes: the Writer PackImpl of the consumer (packer.ts), or the composite as RGBA-f32
enum SinkKind : int {
	SINK_V210 = 0,       // v210.ts:113-195 (the march kernel's only sink)
	SINK_RGBA8 = 1,      // rgba8.ts:69-103   (ScreenConsumer)
	SINK_BGRA8 = 2,      // bgra8.ts:69-103
	SINK_YUV422P10 = 3,  // yuv422p10.ts:126-219: out = Y, out_u, out_v
	SINK_YUV422P8 = 4,   // yuv422p8.ts:126-219 (FFmpegConsumer)
	SINK_YUV420P = 5,    // yuv420p.ts:142-238
	SINK_NV12 = 6,       // nv12.ts:134-240: out = Y, out_u = interleaved chroma
	SINK_RGBA_F32 = 7    // the composite itself as an RGBA-f32 frame (a deferred frame made real: ROUTE payloads, Yadif inputs, host reads)
};

// first pass of a separable Lanczos Transform (pb_march.cu k_lanczos_hpass): every source row of a packed leaf that the
// vertical support of some output line reaches is converted ONCE and filtered horizontally into H (one float4 per output
// column: the three colour sums and the weight sum of the texels inside the image)
struct HPassDesc {
	Leaf lf;             // the packed source leaf with its Lanczos tables and the strip footprints of the consuming launch
	ReadConsts rc;
	ReadK rk;
	LutDesc lut;         // the leaf's gamma table in the one-byte form
	float4 *out;         // H: lf.h rows x xf_w columns
	int xf_w;
	int strip_groups, s0, s1;   // output-column strips as the consuming launch cuts them; [s0, s1] touch the image
	int j_lo, j_hi;      // source rows needed: [j_lo, j_hi)
	uint32_t e_magic, lds_koff;
	int pre_kind;        // 0: the Lanczos first pass above.  1: the interpolated lines of a Yadif leaf (lf.ptr_u / ptr / ptr_v = prev / cur /
	                     // next, lf.yadif = flags) into out: row r of out = line 2 r + (1 - parity) of the field
};

struct FusedDesc {
	int n_layers;
	int out_w, out_h;
	int interlace;     // packer.ts:24-28 Interlace enum: 0, 1 (top), 3 (bottom)
	int out_pitch;     // bytes per output line
	int n_rc;
	void *out;
	void *out_u, *out_v;   // planar sinks
	int sink;              // SinkKind
	int pad_sink_;
	// march kernel
	int strip_groups;  // output groups per strip
	int n_strips;
	int n_luts;        // tables to stage in shared memory (0: gather from the raw tables in global memory)
	int dbg;           // experiment switches (PB_DBG environment variable), 0 in production
	uint32_t e_magic;  // 0x4B000000, handed to the kernel as data so that (w & mask) | e stays one LOP3
	uint32_t lds_koff; // -0x4B000000 (mod 2^32) as data of its own: stays in a uniform register, the operand of LDS.U8 [R + UR] (pb_march.cu LutK::koff)
	int sparse_cm;     // every rc has cm[1] == 0 and cm[10] == 0 (true for all colourMaths YCbCr matrices)
	int any_planar;    // general load path: some leaf is planar 4:2:2 / 4:2:0, or a source width is not a multiple of 6, or the sink is not v210
	int single_strip_groups;   // output groups per strip of the single-layer item loop: 31 stand-alone, 30 as the background pass
	int bg_single;             // the general kernel leaves the background-only strip-pair lines to march_single_items<true>
	unsigned int *bg_counter;   // bg_single: item counter of the second phase, this launch's own (zeroed on the stream before the launch)
	unsigned int bg_base;       // its value when this launch starts: 0
	const unsigned long long *line_pairs;   // bg_single: per output line, bit p set = strip pair p is background-only there
	int single_lines;  // > 0: k_march_single (one v210 layer through an axis-aligned Transform): output lines per work item
	int2 single_strips[64];   // k_march_single: {first source group, source groups} of each 186-px output strip
	int direct_mode;   // one v210 leaf read 1:1 into a v210 sink: the dedicated k_march_direct (192-px strips, all lanes convert)
	int n_t256;        // 1 KiB tables of rgba8 / bgra8 leaves staged behind the row buffers (ReadConsts::t256_slot)
	int big_rows;      // row buffers of 64 source groups (a leaf is scaled down below ~0.47); implies any_planar
	int feat;          // general march variants: abilities this launch needs (pb_march.cu kFeat): 1 Lanczos filtered inside the launch, 2 Yadif leaves, 4 RGBA-f32 sink
	int march_w;       // output pixels the march kernel writes: out_w rounded down to whole v210 groups
	int g_first;       // generic kernel, v210 sink: first output group column to write (the ragged tail after a march launch), else 0
	LutDesc luts[kMaxLuts];   // slot 0 = rc[0]'s table
	LutParams wlp;            // = luts[wc.lut_slot].lp, at a fixed offset for the encoder
	WriteConsts wc;
	ReadConsts rc[kMaxReadConsts];
	ReadK rk[kMaxReadConsts];
	Layer layers[kMaxLayers];
	// march kernel
	int n_ops;
	MarchOp ops[kMaxOps];
	// per strip: bit i (< 24) set if op i can touch the strip (all ops of a transition layer together); bit 24 + l set
	// if layer l is exactly opaque (alpha == 1.0f) on every column of the strip
	uint32_t strip_ops[kMaxStrips];
	const uint32_t *line_ops;         // per output line, same meaning (device memory, out_h entries)
	// Exact occlusion culling: where (strip_ops & line_ops) >> 24 names an opaque layer L, `over` discards everything
	// below it (fma(prev, 1 - 1, l) == l), so ops before layer_first_op[L] are dropped for that strip line.
	int layer_first_op[kMaxLayers];
};

}  // namespace pb
