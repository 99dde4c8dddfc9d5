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
	int jcut;                              // chunks j > jcut certainly start behind the capacity, and so do the tiles >= icut of
	int icut;                              // chunk jcut: they are not coded at all (jcut == nchunks: nothing is cut)
	u32 ref_cut;                           // refinement bits of the chunks < jcut (= tot_ref when nothing was cut)
};

struct EncBuffers {
	const u32 *bs;        // bit-sliced store
	u32 *ent_z, *ent_1, *ent_r; // per (chunk, tile) counts -> exclusive prefixes
	int nent;
	u32 *Z;               // Z[t+1] = zero symbols before token t (mod 2^32); Z[0] = 0
	u32 *signbuf;         // 1 bit per token
	u32 *specbuf;         // 1 bit per token: flush candidate / final token
	u32 *refbuf;          // dense refinement bits, all chunks back to back
	u32 *tile_flags;      // ticket + per token tile: order map ends, resolved end, bit prefix state (zeroed per encode)
	size_t tile_flag_bytes;
	unsigned char *thr_state;  // resolved order at the first token of every thread
	u64 *tile_bitbase;
	EncChunks *chunks;
	EncInfo *info;
	const Sched *sched;   // device copy
	u32 *out;             // zero-initialised output stream words
	u64 out_limit_bits;   // writes are clipped here (buffer keeps 64 bits of slack beyond it)
	u32 max_tokens;
};

int enc_count(const Geom &g, const Sched &hs, const EncBuffers &b, cudaStream_t st, long long *launches);
// prefix_bits: bits of header + root image + plane counts; limit_bits: 8 * capacity, 0 = unlimited.  Work whose first bit
// provably lies behind limit_bits (lower bound: 2 bits per one, 1 per refinement bit; granularity: a tile of 256 groups of
// a chunk) is dropped from the rest of the pipeline -- the reference stops coding at the cap as well (encode.c:193,205,217
// `goto end`, bytes.h:77-78).  enc_token_bound / enc_ref_bound say how many tokens / refinement bits can survive the cut.
int enc_scan_and_setup(const Geom &g, const Sched &hs, const EncBuffers &b, u64 prefix_bits, u64 limit_bits, cudaStream_t st,
                       long long *launches);
u64 enc_token_bound(const Geom &g, const Sched &hs, u64 prefix_bits, u64 limit_bits);
u64 enc_ref_bound(const Geom &g, const Sched &hs, u64 prefix_bits, u64 limit_bits);
int enc_emit(const Geom &g, const Sched &hs, const EncBuffers &b, cudaStream_t st, long long *launches);
// VLI order resolution + token bit lengths; k0 = order after the host-coded prefix
int enc_vli_orders(const EncBuffers &b, int k0, cudaStream_t st, long long *launches);
// scatter tokens and refinement bits into b.out; prefix_bits = bits already occupied (header + root + planes)
int enc_scatter(const Sched &hs, const EncBuffers &b, u64 prefix_bits, u64 tot_ref, cudaStream_t st, long long *launches);

// ---------------------------------------------------------------------------------------------- decoder

#define DWT_DEC_WS 128            // stream slices (64 bits each) per scan window (1 KB of stream)
#define DWT_DEC_MAX_LINEAGE 16    // lineage passes a decode can run
#define DWT_DEC_SUPER 32          // windows per super-window (the resolver's second level)

struct DecState {                 // device-resident cursor of the serial parse (decode.c:187-243); read back by the host
	u64 bitpos;                   // next stream bit
	u64 end_bits;                 // 8 * stream length
	int order;                    // VLI order
	u32 pending;                  // rle.h:66-77 counter `cnt`
	int stopped;                  // EOF / corrupt stream reached: later chunks are skipped
	int level;                    // highest level started
	int missing[48];
	u32 nseg;                     // (chunk, window) segments handed to the emit kernel
	u32 nbulk;                    // super-windows consumed whole: their 32 segments each are written by dec_bulk_kernel
	u32 exact_steps, slow_entries; // resolver statistics: slices stepped token by token, entries into that mode
	u32 n_super, n_window, n_search; // ... super-window rounds, window rounds, end searches
	u32 dbg_hist[10];              // exact slice steps per window visit: 0, 1, 2, <=4, <=8, <=16, <=32, <=64, <=128, visits that joined
	u32 dbg_reason[4];             // entries: chunk start, stray class, pass ends before the join, window not joined
	int guard_tripped;            // the resolver's iteration guard fired (a bug, never a property of a stream)
};

struct DecChunk {                 // per chunk, written by the resolver, read by emit + deposit
	u64 rank_base;                // bit offset of the chunk's member ranks in ones_rank / sign_rank
	u64 ref_bitpos;               // where the refinement bits start
	u32 r0;                       // members covered by the run carried in from earlier chunks
	u32 T;                        // members left for this chunk's own tokens
	int parsed;                   // the significance pass of this chunk ran
	int ref_valid;                // the refinement pass was reached
};

struct DecLink {                  // window w entered with the exit state of class q of window w-1
	u64 mem;                      // members consumed inside the window
	u32 tok;                      // tokens started inside the window
	u32 stray_mem, stray_tok;     // ... of which before the chain joins a canonical chain (saturating)
	unsigned short exit_state;    // state behind the window when the chain never joined (qn == 2)
	unsigned short m;             // slice where it joins (DWT_DEC_WS: never)
	unsigned char qn;             // class behind the window: 0 / 1 canonical, 2 stray, 3 dead
	unsigned char qm;             // canonical class joined at slice m
	unsigned char pad[6];
};

struct DecSeg {                   // one (chunk, window) visit of the true chain
	u32 w;                        // window
	unsigned short j;             // chunk
	unsigned short state;         // entry state (offset | order << 6) at slice i0
	unsigned short i0;            // first slice of the visit
	unsigned short m;             // slice from which the chain is canonical class qm (DWT_DEC_WS: nowhere)
	u32 qm;
	u32 cum0;                     // members of the chunk consumed before the visit (after r0)
	u32 cum_m;                    // ... before slice m
};

struct DecSuper {                 // DWT_DEC_SUPER windows entered with the exit state of class q of the window in front
	u64 mem;                      // members consumed by all of them
	u32 tok;                      // tokens started in them
	unsigned short exit_state;    // state behind the last window when the chain left it without joining (qn == 2)
	unsigned char qn;             // class behind the last window
	unsigned char clean;          // every window exists and is entered with a canonical class: the sums are the whole story
};

struct DecBulk {                  // one super-window consumed whole by chunk j
	u32 w0;                       // first window
	u32 seg_base;                 // its 32 segment records go to seg[seg_base ..]
	u32 cum0;                     // members of the chunk consumed before window w0 (after r0)
	unsigned short j;
	unsigned short cls;           // class the chain enters window w0 with
};

struct DecBuffers {
	u32 *bs;              // bit-sliced store being filled
	u32 *sig;             // significance words [c][GT]
	const u32 *stream;    // stream words (zero padded by >= 64 bytes)
	const u32 *toklut;    // order-0 token table: 4096 length words, then 4096 ones/sign mask words (dec_token_table)
	u64 end_bits;
	u32 nwin;             // scan windows covering the stream
	int in_flight;        // contexts the caller keeps busy on this device (dwt_ctx_set_in_flight): picks the scan kernel
	int scan_mode;        // 0: by in_flight and stream size, 1: parallel scan kernel, 2: serial scan kernel
	u32 *E;               // per slice: canonical entry states of the two classes (e0 | e1 << 16)
	ulonglong2 *P;        // per slice: members consumed by each class since the window start
	u32 *TK;              // per slice: tokens started by each class since the window start (t0 | t1 << 16)
	u32 *winX;            // per window: exit states of the two classes
	u32 *winX2;           // second copy for the lineage passes (they read one and write the other)
	unsigned char *chg;   // 2 * nwin flags: windows whose class-1 chain a lineage pass replaced
	u32 *ext_count;       // DWT_DEC_MAX_LINEAGE counters (zeroed per decode): windows a lineage pass has to walk
	uint2 *ext_list;      // nwin entries: (window, entry state) of those windows
	ulonglong2 *winPT;    // per window: member totals
	u32 *winTT;           // per window: token totals
	DecLink *link;        // [window][class]
	DecSuper *super;      // [super-window][class]
	u32 nsuper;
	DecBulk *bulk;
	DecSeg *seg;
	DecChunk *chunks;
	u32 *tile_sums, *tile_base; // per (channel, tile): (member, refinement) counts and their exclusive prefixes
	u32 *ones_rank, *sign_rank; // rank-space bit vectors, all chunks back to back
	DecState *state;
	const Sched *sched;   // device copy
};

// whole significance/refinement decode of `nchunks` chunks of the schedule into b.bs (zero-initialised by the caller)
int dec_run(const Geom &g, const Sched &hs, const DecBuffers &b, int nchunks, cudaStream_t st, long long *launches);
#define DWT_DEC_LUT_WORDS 8192
void dec_token_table(u32 *host_table); // fills DWT_DEC_LUT_WORDS words
// bits of rank space needed for the first nchunks chunks
u64 dec_rank_bits(const Geom &g, const Sched &hs, int nchunks);
