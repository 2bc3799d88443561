/* snowtri.h -- C ABI of the B200-native multi-camera triangulation engine.
 *
 * This is the drop-in boundary for the one hot path of liaochikon/SnowMocap:
 *     CameraGroup.add_human_2D_points   (reference snowvision/camera.py:234-253)
 *  -> Human_Triangulation               (reference snowvision/triangulation.py:50-93)
 *  -> Human_Triangulation_Condense      (reference snowvision/triangulation.py:95-162)
 * The reference has no FFI of its own (it is pure Python); these entry points are what a
 * ctypes binding inside snowvision would call.
 *
 * Conventions
 *  - Every function returns 0 on success or a negative SNOWTRI_E_* code; no C++ exception
 *    crosses this boundary.  snowtri_last_error() gives a human-readable message.
 *  - All d_* arguments are DEVICE pointers owned by the caller, h_* are HOST pointers.
 *    The library only allocates its own scratch.  Work is enqueued on `stream`
 *    (a cudaStream_t passed as void*, NULL = default stream) and is asynchronous unless noted.
 *  - One handle per device; a handle is not thread-safe.
 *  - Layouts (row-major, innermost last):
 *      kpts    (F, C, P, J, 2) float32   undistorted pixel coordinates (u, v)
 *      scores  (F, C, P, J)    float32   detector confidences
 *      counts  (F, C)          int32     persons present per camera (slots [0,count) valid);
 *                                        NULL means every camera sees P persons
 *      cameras K (C,3,3), R (C,3,3) camera->world, t (C,3) camera centre; float64, host
 *  - There is no CPU fallback: every entry point needs a CUDA device.
 */
#ifndef SNOWTRI_H_
#define SNOWTRI_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct snowtri_handle snowtri_t;

enum {
    SNOWTRI_OK = 0,
    SNOWTRI_E_ARG = -1,        /* bad argument (NULL, out-of-range index, misaligned pointer) */
    SNOWTRI_E_CUDA = -2,       /* a CUDA runtime call failed */
    SNOWTRI_E_UNSUPPORTED = -3,/* problem size not supported by this build */
    SNOWTRI_E_NOMEM = -4
};

/* Compute precision of the fused path (snowtri_run*).  The candidate/condense entry points
 * always compute in float64 like the reference.  The float modes apply to the single-person kernel
 * (one person per camera, shipped thresholds); every other case computes in float64 unless
 * SNOWTRI_PREC_F32_EXPERIMENTAL is selected. */
enum {
    SNOWTRI_PREC_F64 = 0,      /* float64 arithmetic throughout (default; matches the reference) */
    SNOWTRI_PREC_F32 = 1,      /* float32 arithmetic, float64 re-evaluation of decisions near a threshold */
    SNOWTRI_PREC_MIXED = 2,    /* float32 arithmetic, float64 for the ray-distance numerator and the decisions */
    SNOWTRI_PREC_F32_EXPERIMENTAL = 3 /* like F32, and also float32 in the general kernel with several persons per
                                  camera: decisions (nout, membership) stay exact, but joints of wrongly
                                  matched "ghost" clusters are only good to ~1e-3; F32/MIXED use float64 there */
};

/* Camera parameter container on the device; replaces Camera/CameraGroup's K, R, t
 * (reference snowvision/camera.py:16-44, 141-157).  K, R, t are host float64. */
int snowtri_create(snowtri_t** out, int device, int C, const double* K, const double* R, const double* t);
int snowtri_destroy(snowtri_t* h);

/* Thresholds: Human_Triangulation(keypoint_score_threshold, average_score_threshold,
 * distance_threshold) and Human_Triangulation_Condense(condense_distance_tol,
 * condense_person_num_tol, condense_score_tol, center_point_index)
 * (reference snowvision/triangulation.py:50, 95-100; config keys
 * configs/snowmocap_default_config.json:10-16). */
int snowtri_set_params(snowtri_t* h, double kst, double ast, double dthr,
                       double cond_tol, int num_tol, double score_tol, int center);
int snowtri_set_precision(snowtri_t* h, int precision);
/* Launch tuning for tests/benchmarks: frames staged per CTA iteration, CTA cap, block size
 * (256 or 512; -256 = 256 threads with stored rays even for one person per camera); 0 = automatic.
 * One person per camera with the shipped thresholds (ast <= 0, score_tol <= 0, kst >= 0, C <= 8) runs the
 * warp-autonomous single-person kernel, where frames_per_group is the frames per warp tile (<= 32);
 * any non-zero `threads` (-1 = "automatic block size") selects the general fused kernel instead. */
int snowtri_set_tuning(snowtri_t* h, int frames_per_group, int max_ctas, int threads);
/* Several persons per camera, float modes: 2 (default) = second-generation kernels of the streaming general path
 * (matching + centres in one launch, clustering + member decode, clique fuse + person score: 3 launches per chunk);
 * 1 = the first-generation kernels (6 launches), kept for the all-float64 mode and for comparisons.
 * snowtri_last_kernel() says which ran: "general2" (both new kernels), "general2m" (new matching, first-generation
 * fuse: more than 8 cameras) or "general". */
int snowtri_set_general_kernels(snowtri_t* h, int generation);

/* Fused hot path for a batch of F frames: rays -> all camera-pair x person-pair candidates ->
 * gating -> greedy clustering -> score-weighted fuse (main.py:55-71 for every frame).
 *   keypoint_num  Condense's keypoint_num (<= J); output joints per person
 *   Pout          output person slots per frame
 *   d_out         (F, Pout, keypoint_num, 4) float32: x, y, z, keypoint score; unused slots zeroed
 *   d_pscores     (F, Pout) float32 person score; unused slots zeroed
 *   d_nout        (F) int32 persons the reference would emit (if > Pout the frame was truncated)
 */
int snowtri_run(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts,
                int F, int P, int J, int keypoint_num, int Pout,
                float* d_out, float* d_pscores, int* d_nout, void* stream);

/* Same through HOST buffers (pinned memory recommended).  The batch is cut into chunks that flow
 * through three internal streams (ordered after `stream`) -- host->device copies, kernels,
 * device->host copies: chunk i is triangulated and its results travel device->host while the inputs
 * of the chunks behind it travel host->device, so both PCIe directions stay busy and never wait for
 * a kernel.  Returns after everything has arrived.  This is the call a reference-side plugin makes. */
int snowtri_run_host(snowtri_t* h, const float* h_kpts, const float* h_scores, const int* h_counts,
                     int F, int P, int J, int keypoint_num, int Pout,
                     float* h_out, float* h_pscores, int* h_nout, void* stream);

/* Runtime specialisation of the single-person kernel (float modes) for this handle's camera rig and the batch
 * shape: the constants are compiled into the instruction stream with NVRTC (about one second, once per rig and
 * shape; cached in the handle).  mode 0 = never, 1 = automatic (default: batches of >= 65536 frames, or whenever
 * the kernel is already compiled), 2 = always.  If NVRTC or the driver API cannot be loaded, or compilation
 * fails, the precompiled kernel runs instead and snowtri_jit_status() says why. */
int snowtri_set_jit(snowtri_t* h, int mode);
const char* snowtri_jit_status(snowtri_t* h);

/* Frames per chunk of snowtri_run_host's copy/compute pipeline (0 = automatic: about 24 MB of input, and
 * with several persons per camera at least eight frames per SM). */
int snowtri_set_pipeline(snowtri_t* h, int frames_per_chunk);

/* Human_Triangulation alone (reference snowvision/triangulation.py:50-93), float64 results.
 * Candidates are written at their DENSE index n = ((pair*P + pm)*P + ps), pair enumerating
 * (mc < sc) lexicographically, Nc = C*(C-1)/2*P*P per frame:
 *   d_cand (F, Nc, J, 4) float64: x, y, z, gated score
 *   d_avg  (F, Nc) float64 mean score (the reference's person score)
 *   d_keep (F, Nc) int32   1 if the reference would append this candidate, else 0
 * The reference's list order is increasing n over the kept candidates. */
int snowtri_candidates(snowtri_t* h, const float* d_kpts, const float* d_scores, const int* d_counts,
                       int F, int P, int J, double* d_cand, double* d_avg, int* d_keep, void* stream);

/* Human_Triangulation_Condense alone (reference snowvision/triangulation.py:95-162) on
 * caller-supplied candidates, float64:
 *   d_cand (F, N, J, 4) float64: x, y, z, score of candidate i in list order; d_ncand (F) int32 <= N
 *   d_out  (F, Pout, keypoint_num, 4) float64; d_pscores (F, Pout) float64; d_nout (F) int32 */
int snowtri_condense(snowtri_t* h, const double* d_cand, const int* d_ncand, int F, int N, int J,
                     int keypoint_num, int Pout, double* d_out, double* d_pscores, int* d_nout,
                     void* stream);

/* Batched Skew_Ray_Solver (reference snowvision/triangulation.py:24-31): n independent ray pairs.
 * d_hm, d_hs, d_tm, d_ts (n,3) float64 -> d_dist (n), d_mid (n,3). */
int snowtri_skew_ray(snowtri_t* h, int n, const double* d_hm, const double* d_hs,
                     const double* d_tm, const double* d_ts, double* d_dist, double* d_mid, void* stream);

/* Human_Triangulation_Smooth (reference snowvision/triangulation.py:4-22, 164-186; called frame after frame
 * by main.py:72-78): one second-order follower per (person, joint), float64 state kept on the device so a
 * clip can be streamed in batches of consecutive frames.  f, z, r are the reference's smooth_f/z/r.
 *   d_out      (F, Pout, J, 4) float32 (snowtri_run layout) or float64 (snowtri_condense layout): x, y, z are
 *              replaced in place by the smoothed point, the score passes through
 *   d_nout     (F) persons in each frame (as written by snowtri_run / snowtri_condense)
 *   d_nsmooth  (F) persons in the smoothed list: the first frame of a clip passes through unchanged and
 *              fixes n0 = its person count; later frames give min(nout, n0) -- the reference zips persons
 *              with the followers created on the first frame, so later persons are dropped and followers of
 *              absent persons are not advanced.  Slots >= d_nsmooth[f] are left untouched.
 * Batches of up to 256 frames run in one launch with a sequential frame loop (the recurrence is sequential);
 * longer batches are cut into chunks that run in parallel.  When the follower forgets its state within 256 present
 * frames to the last bit of a float64 (the powers of its update matrix fall below 1e-19: 80 frames for the reference's
 * shipped f = 2.5, z = 0.75, r = 0 at 30 fps) every chunk warms up on the frames before it instead of waiting for its
 * predecessor's state: one cooperative launch, the batch read about 1.2 times and written once.  Slower filters take
 * 128-frame chunks with the state handed from chunk to chunk through powers of the (affine) update matrix (a prefix
 * scan of the chunk maps).  snowtri_smooth_set_chunked(s, 0) forces the sequential kernel, (s, 2) the chunk scan,
 * (s, 1) is the default.
 * max_persons bounds n0: create the state with at least as many followers as the FIRST frame of the clip can hold
 * persons (max_persons >= Pout is always enough).  A first frame with more persons than max_persons seeds followers
 * for the first max_persons of them only -- the reference would keep one follower per first-frame person -- and the
 * others are not smoothed on later frames; no error is raised (the person count lives on the device).  The same holds
 * for snowtri_blender_smooth_create. */
typedef struct snowtri_smooth_state snowtri_smooth_t;
int snowtri_smooth_create(snowtri_t* h, snowtri_smooth_t** out, int max_persons, int J, double f, double z, double r);
int snowtri_smooth_destroy(snowtri_smooth_t* s);
int snowtri_smooth_reset(snowtri_t* h, snowtri_smooth_t* s, void* stream);   /* next frame starts a new clip */
int snowtri_smooth_set_chunked(snowtri_smooth_t* s, int enabled);
int snowtri_smooth_run(snowtri_t* h, snowtri_smooth_t* s, float* d_out, const int* d_nout, int* d_nsmooth,
                       int F, int Pout, int J, double delta_time, void* stream);
int snowtri_smooth_run_f64(snowtri_t* h, snowtri_smooth_t* s, double* d_out, const int* d_nout, int* d_nsmooth,
                           int F, int Pout, int J, double delta_time, void* stream);

/* Ragged ingestion: what main.py:50-55 + add_human_2D_points/clear_2D_points (reference
 * snowvision/camera.py:234-261) do per frame, for a whole clip on the device.  The detector's outputs of every
 * (frame, camera) -- (N_fc, J, 2) keypoints and (N_fc, J) scores, persons in detector order -- are
 * concatenated in (frame, camera) order:
 *   d_det_kpts (M, J, 2) float32, d_det_scores (M, J) float32, d_offsets (F*C + 1) int64 CSR row starts
 * and padded into the dense layout snowtri_run consumes: d_kpts (F,C,P,J,2), d_scores (F,C,P,J) (unused slots
 * zeroed), d_counts (F,C) = min(N_fc, P).  Persons beyond the P slots of a camera are dropped. */
int snowtri_pack_ragged(snowtri_t* h, const float* d_det_kpts, const float* d_det_scores,
                        const long long* d_offsets, int F, int P, int J,
                        float* d_kpts, float* d_scores, int* d_counts, void* stream);

/* Optional DLT mode (NOT what the reference computes): one person per camera,
 * every (frame, joint) is triangulated from all cameras whose score passes keypoint_score_threshold by the
 * homogeneous linear method -- rows xn*Q[2]-Q[0], yn*Q[2]-Q[1] with Q = [R^T | -R^T t] and (xn,yn) = K^-1 [u v 1],
 * solution = eigenvector of the 4x4 normal matrix with the smallest eigenvalue.
 *   d_kpts (F,C,1,J,2), d_scores (F,C,1,J) float32;  d_out (F,J,4) float32: x, y, z, number of views used
 *   (fewer than two views: zeros).  accumulate_f64 = 1 builds the normal matrix in float64 instead of float32. */
int snowtri_dlt_run(snowtri_t* h, const float* d_kpts, const float* d_scores, int F, int J, float* d_out,
                    int accumulate_f64, void* stream);

/* Blender control points (reference snowvision/blender.py; main.py:80-87 runs both steps on every frame).
 *
 * snowtri_blender_run = Human_Triangulation_Blender (blender.py:93-143, helpers :11-96): every person row of the
 * snowtri_run / snowtri_condense output gives the 24 control points of configs/blender_armature_profile.json, in
 * that file's order (0 root_position, 1 root_rotation, 2 clavicle_r_ik, 3 clavicle_l_ik, 4 arm_r_ik, 5 arm_r_pole,
 * 6 arm_l_ik, 7 arm_l_pole, 8 leg_r_ik, 9 leg_r_pole, 10 leg_l_ik, 11 leg_l_pole, 12 hand_r_ik, 13 hand_r_pole,
 * 14 hand_l_ik, 15 hand_l_pole, 16 foot_r_ik, 17 foot_r_pole, 18 foot_l_ik, 19 foot_l_pole, 20 chest_ik,
 * 21 chest_pole, 22 head_ik, 23 head_pole):
 *   d_points (F, Pout, J, 4) float32 (snowtri_run) or float64 (snowtri_condense): x, y, z, score; J >= 130
 *   d_nout   (F) persons per frame, or NULL = every row holds a person; rows >= d_nout[f] give zeros / valid 0
 *   d_ctrl   (F, Pout, 24, 4): (x, y, z, 0) for positions, (w, x, y, z) for root_rotation (blender.py:27)
 *   d_valid  (F, Pout) uint32: bit k set = control point k has no NaN component = the reference's score 1
 *            (blender.py:135-138).  The reference raises LinAlgError from SciPy on a NaN root rotation; the
 *            batch call writes NaN and clears bit 1 instead.
 * The arithmetic is float64 for both layouts. */
int snowtri_blender_run(snowtri_t* h, const float* d_points, const int* d_nout, int F, int Pout, int J,
                        float* d_ctrl, unsigned* d_valid, void* stream);
int snowtri_blender_run_f64(snowtri_t* h, const double* d_points, const int* d_nout, int F, int Pout, int J,
                            double* d_ctrl, unsigned* d_valid, void* stream);

/* Human_Triangulation_Blender_Smooth (blender.py:145-178): one second-order follower (triangulation.py:4-22) per
 * (person, control point), float64 state on the device so a clip can be streamed in batches of consecutive
 * frames.  fzr (24,3) HOST array: f, z, r of every control point (configs/blender_smooth_profile.json).
 * First frame of a clip: followers start at the control point (at zero where it is invalid), the frame passes
 * through.  Later frames: persons are zipped by position with the first frame's followers (d_nsmooth[f] =
 * min(d_nout[f], first frame's count)); a control point whose valid bit is 0 re-feeds the follower its previous
 * input (blender.py:159-160).  d_ctrl is smoothed in place; d_valid passes through untouched. */
typedef struct snowtri_blender_smooth_state snowtri_blender_smooth_t;
int snowtri_blender_smooth_create(snowtri_t* h, snowtri_blender_smooth_t** out, int max_persons, const double* fzr);
int snowtri_blender_smooth_destroy(snowtri_blender_smooth_t* s);
int snowtri_blender_smooth_reset(snowtri_t* h, snowtri_blender_smooth_t* s, void* stream);
/* Batches of more than 256 frames are cut into 128-frame chunks that run in parallel (every chunk is an affine map
 * of its start state; the maps are chained by a short sequential pass); enabled = 0 forces the sequential kernel. */
int snowtri_blender_smooth_set_chunked(snowtri_blender_smooth_t* s, int enabled);
int snowtri_blender_smooth_run(snowtri_t* h, snowtri_blender_smooth_t* s, float* d_ctrl, const unsigned* d_valid,
                               const int* d_nout, int* d_nsmooth, int F, int Pout, double delta_time, void* stream);
int snowtri_blender_smooth_run_f64(snowtri_t* h, snowtri_blender_smooth_t* s, double* d_ctrl, const unsigned* d_valid,
                                   const int* d_nout, int* d_nsmooth, int F, int Pout, double delta_time,
                                   void* stream);

/* The per-frame body of the reference's main.py (:55-87) for a whole clip in one call: snowtri_run (keypoint_num = J)
 * -> snowtri_smooth_run (if sm != NULL) -> snowtri_blender_run -> snowtri_blender_smooth_run (if bs != NULL), back to
 * back on `stream`; arguments as documented at those entry points.  d_nsmooth is required with sm, d_nfinal with bs;
 * the control points are derived from the persons counted by d_nsmooth (or d_nout without sm). */
int snowtri_clip_run(snowtri_t* h, snowtri_smooth_t* sm, snowtri_blender_smooth_t* bs, const float* d_kpts,
                     const float* d_scores, const int* d_counts, int F, int P, int J, int Pout, float* d_out,
                     float* d_pscores, int* d_nout, int* d_nsmooth, float* d_ctrl, unsigned* d_valid, int* d_nfinal,
                     double delta_time, void* stream);

/* Final all-gather of the 3D joints across the GPUs of one box (north_star: frames shard across the GPUs, "NCCL
 * over NVLink appears only as a final all-gather of 3D joints").  One process per GPU, one handle per process.
 * NCCL ("libnccl.so.2") is loaded with dlopen at first use.  Either pass the host's own ncclComm_t, or let the handle
 * own one: rank 0 fills 128 bytes with snowtri_comm_unique_id, the host ships them to every rank, every rank calls
 * snowtri_comm_init(h, id, nranks, rank) (collective).
 *   d_send  this rank's block (bytes_per_rank bytes, e.g. its (F/N, Pout, J, 4) float32 output)
 *   d_recv  nranks * bytes_per_rank bytes, rank r's block at offset r * bytes_per_rank (frame order when ranks own
 *           consecutive frame blocks of equal size)
 * Asynchronous on `stream`. */
int snowtri_comm_unique_id(void* id128);
int snowtri_comm_init(snowtri_t* h, const void* id128, int nranks, int rank);
int snowtri_comm_destroy(snowtri_t* h);
int snowtri_allgather(snowtri_t* h, const void* d_send, void* d_recv, size_t bytes_per_rank, void* nccl_comm_or_null,
                      void* stream);

/* Introspection. */
const char* snowtri_last_error(snowtri_t* h);       /* also valid with h == NULL (create failures) */
long long snowtri_launch_count(snowtri_t* h);       /* kernels launched through this handle so far */
int snowtri_last_launch_info(snowtri_t* h, int* grid, int* block, int* smem_bytes, int* frames_per_group);
const char* snowtri_last_kernel(snowtri_t* h);      /* "p1", "p1-jit", "general", "general2", "general2m", "fused" or "fused-fly"; "" before any run */
int snowtri_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SNOWTRI_H_ */
