// fsg_internal.cuh — shared declarations of libfsg's translation units (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/fsg.h"

// ---------------------------------------------------------------------------------------------
// Device-side constants of one context.  The double members are the reference's macros
// (FluidGPU.cuh:1-31); the float members are the thresholds / coefficients the fp32 pair path
// uses, derived on the host so that every comparison decides exactly like the reference's
// double comparison of a float operand.
// ---------------------------------------------------------------------------------------------
struct FsgDev {
    int G, G2, numcells;
    int x0, x1;      // bin layers this context owns (0, G without slab decomposition)
    int rl, rr;      // slab contexts: layers [x0, rl) and [rr, x1) are within two layers of a face that has a neighbour
    int bx0, bx1;    // interior layers [bx0, bx1): home bins outside are the slab's boundary bins (done first when the
                     // exchange overlaps the interior; == x0, x1 otherwise)
    int uni_open;    // unidyn: 1 = no unit-box floor / walls in Particle::update (fsg_config.unidyn_open_box)
    int sym;         // 1: the symmetric pair kernel runs (fsg_pair_v3.cu): the ghost layer x0 - 1 is listed as home bins too
    int dead;        // key of a slot that no longer holds a particle of this slab (sorts last, is trimmed)
    int kx0, kx1;    // sorted-ghost slab pipeline (fsg_slab2.cu): bin ids x0*G^2 and x1*G^2, where the particle array is NOT contiguous
                     // (ghost layers live in their own zones); INT_MIN otherwise
    int cap, bin_cap;
    float origin;
    double cellsize, h, dt, gravity, sound, alpha_fluid, alpha_boundary;
    // fp32 pair constants
    float d2_max;    // largest d2 with sqrtf(d2) <= 2h (as the double compare FluidGPU.cu:236 decides)
    float d2_h;      // largest d2 with sqrtf(d2) <= h_le
    float h_le;      // largest float r with (double)r <= h        FluidGPU.cu:12
    float h_lt;      // largest float r with (double)r <  h        FluidGPU.cu:36
    float twoh_lt;   // largest float r with (double)r <  2h       FluidGPU.cu:15
    float hf;        // (float)h
    float inv_h;     // 1/h
    float w_c;       // 1/3.14159/powf(h,3)                        FluidGPU.cu:13
    float dw_c;      // -45/3.14159/powf(h,6)                      FluidGPU.cu:37
    float w0;        // kernel(0)                                  FluidGPU.cuh:166
    float eps;       // 0.01*powf(h,2)                             FluidGPU.cu:255
    float visc_c;    // ALPHA_FLUID*SOUND
    float visc_q;    // 50*1.0/SOUND
};

// SoA particle state: four float4 streams = 64 B per particle (DESIGN.md "Data layout").
//   posd = (x, y, z, dens)   — sign bit of dens carries Particle::boundary
//   velp = (vx, vy, vz, press)
//   accf = (ax, ay, az, flags bits: bit0 boundary, bit1 solid)
//   dpi  = (delpressx, delpressy, delpressz, index bits)
struct FsgState {
    float4 *posd, *velp, *accf, *dpi;
    float4 *mix;     // unidyn model only: (solid, fluid, -, -)   FluidGPU-unidyn.cuh:180-181
    float *stress;   // unidyn model, single-device contexts: [n][18] stress_tensor[3][3] then stress_rate[3][3] (granular scenes)
};

struct FsgFrameWriter;        // fsg_frame.cu: copy stream, pinned staging slots and the writer thread of the asynchronous frame output

struct fsg_ctx {
    fsg_config cfg;
    FsgDev dev;
    int device;
    cudaStream_t stream;
    bool own_stream;
    std::string err;

    int64_t cap;        // particle capacity
    int64_t n;          // particle slots in use (slab contexts: may include dead slots until the next sort)
    int64_t n_sorted;   // length of the key array the bin tables were built from
    FsgState A, B;      // A: sorted pre-update state of the last step; B: post-update state
    float4 *carryB, *carryA;   // accumulators carried into the first step after upload (newdens, newdelpress xyz)
    bool carry_live;
    // Deferred update (single-device base contexts, uncapped fp32 pair kernels): after a step the post-update state B is NOT
    // materialised — the state is (A, sums [, carryA]) and keysB already holds the bin ids after the update (predicted_key);
    // the next step's reorder applies the update while it gathers.  fsg_materialize() runs k_update when somebody needs B.
    bool deferred;          // B is stale: (A, sums) is the state
    bool carry_pending;     // carryA still has to be added to the sums of the deferred update (first step after an upload)
    int defer_mode;         // -1 not decided, 0 off, 1 on (FSG_DEFER_UPDATE)
    float4 *sums;              // pair sums of the current step (newdens, newdelpress xyz), sorted order
    float4 *sums2;             // unidyn: (diffusion xyz, delfluid)
    float *mixA, *mixB;        // unidyn, mixed-phase scenes: pass-A / pass-B pair sums per sorted slot (fsg_unidyn_mixed.cu)
    bool mixed;                // the uploaded unidyn scene has a non-boundary particle with solid != 0
    float *vizb;               // unidyn: |diffusion|^2 of the last step (mykernel2's b3, FluidGPU-unidyn.cu:466)
    int upload_flags;          // host copy of the flags the upload kernels raise
    bool has_boundary;         // any Particle::boundary set in the uploaded scene
    int *keysB;         // bin ids belonging to B (order of B)
    int *keysA;         // sorted bin ids (order of A)
    int *perm, *iota;
    int *start, *end;   // dense bin tables, -1 = empty  (FluidGPU.cu:106-117)
    int *binlist[2];    // ids of the occupied home bins (unordered), ping-pong
    int *counters;      // [0..1] nocc ping-pong, [2] work counter, [3] n_live, [4] any-boundary flag, [5] n_keep,
                        // [14] the sorted key array was found out of order by k_reorder (never expected; reported by fsg_get_stats / downloads),
                        // [16..19] slab contexts: sorted-slot bounds found by k_reorder — first slot of layer rl, of layer rr (the pack's
                        // two-layer regions), of layer x0 + 1, of layer x1 - 1 (the face layers the sorted-ghost pipeline sends),
                        // [20..21] blocks of k_slab2_ghost_send that have finished (left, right)
    unsigned long long *dstats;   // [0] tested, [1] in range, [2] dropped
    int *slab_cnt;      // slab pack: per-warp counts of the 4 message categories, then their exclusive scan
    void *scan_tmp;
    size_t scan_tmp_bytes;
    int64_t slab_warps;
    void *outbox[2];    // slab messages packed here (to left, to right) when the library owns the buffers
    void *inbox[4];     // [2*side + parity]: from left / from right, double-buffered by step parity
    void *peer_inbox[4];// the neighbours' inboxes mapped through CUDA IPC: [0..1] left neighbour's from-right, [2..3] right neighbour's from-left
    int64_t msg_cap_m, msg_cap_g;
    long long seq_send, seq_recv;   // exchange sequence numbers (stamps at the tail of every message)
    volatile int *host_flag;        // pinned, device-mapped: raised by k_slab_wait when a neighbour's message never arrived (sticky)
    int *host_flag_dev;
    unsigned long long slab_timeout_ns;
    // Sorted-ghost pipeline (fsg_slab2.cu; base model, library-owned peer messages): ghosts never pass through the sort.  The top
    // 2 * msg_cap_g slots of every particle array are the two ghost zones, the context works on n_own = cap - 2 * msg_cap_g slots,
    // the update is deferred like on a single device, migrants travel with their pending pair sums before the sort and the face
    // layers of the SORTED state are copied into the neighbours' ghost zones after the reorder.
    bool slab2;
    bool slab2_split;   // in-process slab groups: fsg_step stops after the ghost send, fsg_slab_step_finish does the rest
    bool step_pending, pending_prof, ghost_prof;
    int pending_nxt;
    bool slab2_mid;     // between fsg_slab_pack_send and fsg_step: migrants have left, B cannot be materialised
    int64_t n_own;
    int gh_par_last;    // parity of the ghost messages whose bins are in the tables (-1: none)
    long long seq_ghost;
    size_t mig_bytes, gh_bytes;      // the two parts of a library-owned message buffer in this mode
    std::vector<cudaEvent_t> ev_ghost;   // profiling: pairs of events around the ghost exchange of every step
    double ghost_ms;
    int64_t ghost_steps;
    bool keep_foreign;  // slab contexts: uploads are not filtered by position (fsg_slab_keep_foreign)
    bool peer_local;    // peer_inbox holds plain pointers of this process (fsg_slab_set_peer), not IPC mappings
    bool overlap;       // pack + copies of the NEXT step's messages run on `comm` behind the boundary bins, beside the interior bins
    bool sent_ahead;    // the messages of the next step have already been issued by fsg_step
    cudaStream_t comm;
    cudaEvent_t ev_boundary, ev_sent;
    int *binlistB;      // boundary home bins of the current step (overlap mode)
    void *sort_tmp;
    // nearly-sorted key sort (fsg_nsort.cu): third key buffer (the sort reads keysB + keysA and writes keysC, then A <-> C), workspace
    int *keysC;
    void *ns_ws;
    size_t ns_ws_bytes;
    int ns_mode;        // -1 not decided yet, 0 off, 1 on (FSG_SORT_MERGE)
    bool keys_prev_valid;   // keysA holds the sorted keys of the step that produced B / keysB (same slot order)
    int64_t ns_used;
    // unidyn particle merging / splitting (fsg_unidyn_adapt.cu)
    int *adapt_ws;          // nn | split flags | their scan: three int arrays of `cap`, then 4 counters
    void *adapt_scan;
    size_t adapt_scan_bytes;
    int adapt_next_index;   // Particle::index of the next child
    int64_t adapt_counts[3], adapt_totals[3];
    FsgFrameWriter *frame_writer;   // created by the first fsg_write_frame_async
    void *stage;        // device staging area for host<->device conversion
    size_t stage_bytes;
    size_t sort_tmp_bytes;
    int sort_bits;
    int cur;            // which binlist/nocc slot the last step used
    bool tables_dirty;  // start/end hold the last step's entries
    int64_t steps;
    int64_t launches;
    int sm_count;
    bool profiling;     // record events around the phases of every step (fsg_set_profiling)
    std::vector<cudaEvent_t> ev_pool, ev_used;   // 5 events per profiled step
    double phase_ms[4];
    int64_t phase_steps;
};

cudaError_t fsg_sort_pairs_int(void *tmp, size_t tmp_bytes, const int *keys_in, int *keys_out, const int *vals_in,
                               int *vals_out, int64_t n, cudaStream_t s);
size_t fsg_sort_int_temp_bytes(int64_t n);
// fsg_sort.cu — stable radix sort of (bin id, slot) pairs; the reference's thrust::sort_by_key (solver.cu:181)
size_t fsg_sort_temp_bytes(int64_t n, int bits);
cudaError_t fsg_sort_pairs(void *tmp, size_t tmp_bytes, const int *keys_in, int *keys_out, const int *vals_in,
                           int *vals_out, int64_t n, int bits, cudaStream_t s);
// fsg_nsort.cu — the same sort for an almost sorted key array (hand-written: flag movers / scan / radix sort of the movers / merge by
// ranking), no host synchronisation.  keys_out / perm_out must not alias the inputs.
size_t fsg_nsort_bytes(int64_t n);
cudaError_t fsg_nsort(void *ws, const int *keys_new, const int *keys_prev, int *keys_out, int *perm_out, int64_t n, int bits, int sm_count,
                      cudaStream_t s, int *launches);

// fsg_base_kernels.cu
cudaError_t fsg_launch_iota(int *p, int64_t n, cudaStream_t s);
cudaError_t fsg_launch_fill(int *p, int v, int64_t n, cudaStream_t s);
cudaError_t fsg_launch_step_counters(int *counters, int nxt, int n, bool slab, unsigned long long *dstats, cudaStream_t s);
cudaError_t fsg_launch_keys(const FsgDev &d, const float4 *posd, int *keys, int64_t n, int *any_boundary, bool slab_filter,
                            const int *slot_state, cudaStream_t s);
cudaError_t fsg_launch_reset_tables_keys(const FsgDev &d, const int *keysA, int *start, int *end, int64_t n, cudaStream_t s);
// fsg_slab.cu
size_t fsg_scan_temp_bytes(int64_t n);
cudaError_t fsg_scan_exclusive(void *tmp, size_t tmp_bytes, const int *in, int *out, int64_t n, cudaStream_t s);
cudaError_t fsg_launch_reset_tables(const int *binlist, const int *nocc, const int *keysA, int *start, int *end,
                                    int64_t n, cudaStream_t s);
cudaError_t fsg_launch_reorder(const FsgDev &d, int64_t n, const int *perm, const int *keysA, FsgState src,
                               FsgState dst, const float4 *carry_src, float4 *carry_dst, const float4 *sums_src, int *keys_next, int *start,
                               int *end, int *binlist, int *nocc, int *binlistB, int *noccB, int *nlive, int *nkeep, int *ranges, int *order_flag,
                               cudaStream_t s);
int fsg_slab_sticky_error(fsg_ctx *c); // fsg_slab.cu: FSG_E_STATE once a device-side wait for a neighbour has timed out
int fsg_slab_send_next(fsg_ctx *c);   // fsg_slab.cu: pack + copies of the next step's messages on c->comm (overlap mode)
// fsg_slab2.cu — the sorted-ghost slab pipeline
int fsg_slab2_engage(fsg_ctx *c, int64_t cap_m, int64_t cap_g, size_t *bytes);   // decides the mode at fsg_slab_alloc_messages; message buffer size
int fsg_slab2_pack_send(fsg_ctx *c);         // migrants (pre-update state + pending sums) -> the neighbours' inboxes
int fsg_slab2_unpack_recv(fsg_ctx *c);       // waits for the neighbours' migrants, appends them behind the slots in use
int fsg_slab2_ghost_send(fsg_ctx *c);            // after the reorder: the sorted face layers -> the neighbours' ghost messages
int fsg_slab2_ghost_recv(fsg_ctx *c, int nxt);   // wait for the neighbours' ghosts, install them + their bins
int fsg_slab2_reset_ghost_tables(fsg_ctx *c);        // start / end = -1 for the ghost bins of the last step
cudaError_t fsg_launch_pair_update(const fsg_ctx *c, int64_t n, const int *binlist, const int *nocc, int *work,
                                   const float4 *carry, int *launches, cudaStream_t s);
cudaError_t fsg_launch_unpack_aos(int model, const unsigned char *aos, int64_t n, FsgState st, float4 *carry,
                                  cudaStream_t s);
cudaError_t fsg_launch_pack_aos(int model, unsigned char *aos, int64_t n, FsgState st, const float4 *carry,
                                const int *keys, float p0, cudaStream_t s);
cudaError_t fsg_launch_export_viz(int64_t n, const float4 *posd, const int *keys, float *spts, float *a3, float *b3,
                                  cudaStream_t s);
cudaError_t fsg_launch_plume(const FsgDev &d, double spacing, double jitter, uint64_t seed, double gravity,
                             FsgState st, float4 *carry, int64_t capacity, unsigned long long *count, cudaStream_t s);

// fsg_unidyn.cu
cudaError_t fsg_launch_unidyn(fsg_ctx *c, int64_t n, const int *binlist, const int *nocc, int *work, const float4 *carry,
                              int *launches, cudaStream_t s);
// fsg_unidyn_adapt.cu
cudaError_t fsg_unidyn_adapt_pre(fsg_ctx *c, int64_t n, cudaStream_t s);     // merge + split marks, between the pair sums and the update
int fsg_unidyn_adapt_post(fsg_ctx *c, int64_t n);                            // children appended; c->n grows (one read-back)
cudaError_t fsg_launch_export_mass(int64_t n, const float4 *mix, float *a3, cudaStream_t s);
cudaError_t fsg_launch_split_table(const fsg_ctx *c, int *split, cudaStream_t s);
cudaError_t fsg_launch_unpack_aos_unidyn(const unsigned char *aos, int64_t n, FsgState st, float4 *carry, int *bad, cudaStream_t s);
cudaError_t fsg_launch_pack_aos_unidyn(unsigned char *aos, int64_t n, FsgState st, const float4 *carry, const int *keys, const FsgDev &d,
                                       cudaStream_t s);

void fsg_frame_writer_destroy(fsg_ctx *c);   // fsg_frame.cu: writes the pending frames, stops the writer thread, frees its buffers
int fsg_materialize(fsg_ctx *c);           // fsg_api.cu: makes B / keysB the post-update state (no-op unless the update is deferred)
void fsg_derive_constants(const fsg_config &cfg, FsgDev &d);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: a process that creates contexts on
// several devices has to opt every kernel in on each of them.  One FsgAttrOnce per launcher; `need()` is true the first time
// the CURRENT device is seen (lock-free, any thread).
#include <atomic>
struct FsgAttrOnce {
    std::atomic<unsigned long long> seen[4];       // bit per device ordinal, 256 devices
    bool need()
    {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 256) return true;
        const unsigned long long bit = 1ull << (dev & 63);
        return (seen[dev >> 6].fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
    }
};
void fsg_update_pair_mode(fsg_ctx *c);     // sets c->dev.sym from the configuration and the overlap / stats switches
