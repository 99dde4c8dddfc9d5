// coder.cuh -- interface of the parallel bit-plane coder (coder_enc.cu) and decoder (coder_dec.cu)
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------- encoder

// Device-resident working set of one encode.  All arrays are owned by the context (pipeline.cu).
struct EncChunks {                         // filled by enc_chunk_setup (device)
	u32 tok_start[DWT_MAX_CHUNKS + 1];     // index of chunk j's first token; [nchunks] = index of the final token
	u32 ref_start[DWT_MAX_CHUNKS + 1];     // dense refinement-bit offset of chunk j
	u32 tok_adj[DWT_MAX_CHUNKS + 1];       // flush candidates before chunk j
	u32 nref[DWT_MAX_CHUNKS];
	u64 ref_pos[DWT_MAX_CHUNKS];           // stream bit position of chunk j's refinement block (set by scatter)
};

struct EncInfo {                           // read back by the host
	u64 tot_zero, tot_one, tot_ref;
	u32 ntok, ntiles;
	u64 tok_bits;                          // bits of all tokens (VLI + sign)
	u64 total_bits;                        // header + prefix + payload, before padding
	u32 opaque_tiles;
	int error;
};

struct EncBuffers {
	const u32 *bs;        // bit-sliced store
	u32 *ent_z, *ent_1, *ent_r; // per (chunk, tile) counts -> exclusive prefixes
	int nent;
	u32 *Z;               // Z[t+1] = zero symbols before token t (mod 2^32); Z[0] = 0
	u32 *signbuf;         // 1 bit per token
	u32 *specbuf;         // 1 bit per token: flush candidate / final token
	u32 *refbuf;          // dense refinement bits, all chunks back to back
	u32 *tile_lo, *tile_hi, *tile_start; // VLI order at tile end for entry order 0 / 31; resolved entry order
	unsigned char *thr_state;  // resolved order at the first token of every thread
	u32 *tile_bits;
	u64 *tile_bitbase;
	EncChunks *chunks;
	EncInfo *info;
	const Sched *sched;   // device copy
	u32 *out;             // zero-initialised output stream words
	u64 out_limit_bits;   // writes are clipped here (buffer keeps 64 bits of slack beyond it)
	u32 max_tokens;
};

int enc_count(const Geom &g, const Sched &hs, const EncBuffers &b, cudaStream_t st, long long *launches);
int enc_scan_and_setup(const Geom &g, const Sched &hs, const EncBuffers &b, cudaStream_t st, long long *launches);
int enc_emit(const Geom &g, const Sched &hs, const EncBuffers &b, cudaStream_t st, long long *launches);
// VLI order resolution + token bit lengths; k0 = order after the host-coded prefix
int enc_vli_orders(const EncBuffers &b, int k0, cudaStream_t st, long long *launches);
// scatter tokens and refinement bits into b.out; prefix_bits = bits already occupied (header + root + planes)
int enc_scatter(const Sched &hs, const EncBuffers &b, u64 prefix_bits, u64 tot_ref, cudaStream_t st, long long *launches);

// ---------------------------------------------------------------------------------------------- decoder

struct DecState {                 // device-resident cursor of the serial parse (decode.c:187-243)
	u64 bitpos;                   // next stream bit
	u64 end_bits;                 // 8 * stream length
	u64 ref_bitpos;               // where the current chunk's refinement bits start
	int order;                    // VLI order
	u32 pending;                  // rle.h:66-77 counter `cnt`
	int stopped;                  // EOF / corrupt stream reached: later chunks are skipped
	int level;                    // highest level started
	int missing[48];
	u32 n_member, n_ref;          // per-chunk totals (scratch)
	int ref_valid;                // the refinement pass of the current chunk was reached
	int chunk_done;               // number of chunks whose parse ran
	u64 c_bitpos;                 // snapshot of (bitpos, order, pending) at the start of the current chunk
	int c_order;
	u32 c_pending;
	u32 ticket;                   // parse windows handed out so far (reset per chunk)
	u32 published;                // parse windows whose exit state is known
	int done;                     // the chunk's significance pass has found its end
	u32 dbg_windows, dbg_iters, dbg_walk; // parse statistics: windows up to the end, fix-up iterations, walk steps
	u64 dbg_cyc[4];               // SM cycles of thread 0: window set-up, wait for the predecessor, walk
};

struct DecBuffers {
	u32 *bs;              // bit-sliced store being filled
	u32 *sig;             // significance words [c][GT]
	const u32 *stream;    // stream words (zero padded by >= 64 bytes)
	u32 *tile_sums, *tile_base; // per tile (member, refinement) counts and their exclusive prefixes (one level)
	u32 *ones_rank, *sign_rank; // rank-space bit vectors of the current chunk (adjacent: one memset)
	u64 *win_state, *win_rank;  // per parse window: published exit state / inclusive member count
	int nwin_cap;
	int parse_ctas;       // persistent parse CTAs (one per SM)
	DecState *state;
};

int dec_setup(void);  // one-time kernel attribute setup
int dec_chunk(const Geom &g, const Sched &hs, const DecBuffers &b, int j, cudaStream_t st, long long *launches);
