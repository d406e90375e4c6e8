/*
 * alrender.h — C-ABI of the B200-native renderer for AudibleLight's synthesis hot path.
 *
 * The reference (AudibleLight v0.1.2, pure Python) has no FFI for this path; its seam is three module-level
 * functions of audiblelight/synthesize.py that Scene.generate imports at call time (core.py:1828-1838):
 *
 *     render_event_audio(event, irs, mic_alias, ref_db, ...)        synthesize.py:507-608
 *     render_audio_for_all_scene_events(scene, ignore_cache)        synthesize.py:613-677
 *     generate_scene_audio_from_events(scene)                       synthesize.py:314-401
 *
 * The Python host layer (audiblelight_b200/synthesize.py) keeps those names and signatures, does the
 * validation / exception mapping / bit-exact timing arithmetic, and hands flat descriptors to the entry
 * points below through ctypes.  Everything here is plain pointers and sizes; no torch
 * or C++ types cross the boundary.  All functions return ALR_OK (0) or a negative status;
 * alr_last_error() gives the message of the last failure on the calling thread.
 *
 * One context per GPU.  A context owns only its workspace (twiddle tables, spectra, descriptors, staging);
 * every input and output buffer belongs to the caller.
 */
#ifndef ALRENDER_H
#define ALRENDER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALR_VERSION 200

enum {
  ALR_OK = 0,
  ALR_ERR_INVALID = -1,   /* bad argument / descriptor */
  ALR_ERR_CUDA = -2,      /* CUDA runtime failure (message has the CUDA error string) */
  ALR_ERR_NOMEM = -3,
  ALR_ERR_NO_DEVICE = -4  /* no usable CUDA device: there is NO CPU fallback */
};

/* where the caller's buffers live */
enum { ALR_MEM_HOST = 0, ALR_MEM_DEVICE = 1 };

/* what happens to the convolved signal y of an event after the convolution */
enum {
  /* reference render_event_audio: y * snr / max|y| then * 10^((ref_db+snr)/20) / (mean|.| + tiny)
     (apply_snr synthesize.py:40-49 + db_to_multiplier :52-68, used at :594-599) */
  ALR_GAIN_EVENT = 0,
  /* raw convolution output, as time_invariant_convolution (:71-106) / time_variant_convolution (:277-310) */
  ALR_GAIN_NONE = 1
};

typedef struct alr_context alr_context;

/*
 * Linear event augmentations applied to the dry audio BEFORE the convolution ("next" row f1 of the scope table):
 * the effects of audiblelight/augmentation.py that are linear filters, as Event.load_audio applies them
 * (event.py:530-532), followed by its peak normalisation (event.py:535-536) when alr_event.normalize_audio is set.
 * Fade / Invert / Reverse are fully specified by the reference's own numpy code (augmentation.py:1490-1601) and are
 * pinned by golden vectors; Gain and the IIR filters wrap pedalboard (JUCE) and librosa, which are neither vendored
 * nor installed offline: their arithmetic here follows the published formulas (host side: audiblelight_b200/
 * augment.py) and is checked against scipy.signal.lfilter only -> parity UNPINNED for those.
 */
enum {
  ALR_AUG_GAIN = 0,        /* y = p[0] * x                         (Gain, augmentation.py:1105; p[0] = 10^(dB/20)) */
  ALR_AUG_INVERT = 1,      /* y = -x                               (Invert, :1557) */
  ALR_AUG_REVERSE = 2,     /* y[n] = x[L-1-n]                      (Reverse, :1583) */
  ALR_AUG_FADE = 3,        /* y = x * fade_in * fade_out           (Fade, :1403-1554) */
  ALR_AUG_BIQUAD = 4,      /* y[n] = p0 x[n] + p1 x[n-1] + p2 x[n-2] - p3 y[n-1] - p4 y[n-2], zero initial state
                              (Low/HighpassFilter :303/:406 first order; Low/HighShelfFilter :348/:449; PeakFilter :643) */
  ALR_AUG_PREEMPHASIS = 5, /* librosa.effects.preemphasis(coef = p[0])   (:1350) */
  ALR_AUG_DEEMPHASIS = 6,  /* librosa.effects.deemphasis(coef = p[0])    (:1388) */
  ALR_AUG_DELAY = 7        /* feedback delay (Delay, :1046, pedalboard.Delay): D = p[0] = int(delay_seconds * sr) samples,
                              d[n] = x[n-D] + p[1] d[n-D] (p[1] = feedback), y[n] = (1 - p[2]) x[n] + p[2] d[n] (p[2] = mix);
                              D == 0: y = x */
};
enum { ALR_FADE_LINEAR = 0, ALR_FADE_EXPONENTIAL = 1, ALR_FADE_LOGARITHMIC = 2, ALR_FADE_QUARTER_SINE = 3,
       ALR_FADE_HALF_SINE = 4, ALR_FADE_NONE = 5 };

typedef struct alr_aug_op {
  int32_t type;             /* ALR_AUG_* */
  int32_t fade_in_shape;    /* ALR_FADE_* */
  int32_t fade_out_shape;
  int32_t fade_in_samples;  /* min(int(round(fade_in_len * sr)), L)  (augmentation.py:1536-1541) */
  int32_t fade_out_samples;
  int32_t reserved;
  double p[6];
} alr_aug_op;

/*
 * One (event, microphone) render = one call of the reference's render_event_audio (synthesize.py:507).
 * All sample data is float32 (the reference computes in float64; the contract is max-abs error <= 1e-5 of
 * full scale, BASELINE.json north_star).
 */
typedef struct alr_event {
  /* ---- inputs ---- */
  const float* audio;       /* (n_audio) mono dry audio == Event.load_audio() (event.py:496-539) */
  int64_t n_audio;          /* Lx */
  const float* irs;         /* RIR taps; element (capsule c, emitter l, tap t) at irs[c*ir_stride_c + l*ir_stride_n + t]
                               (the reference layout is (C, N, Lh), worldstate.py:2183-2255) */
  int64_t ir_stride_c;      /* in elements */
  int64_t ir_stride_n;      /* in elements */
  int32_t n_channels;       /* C: capsules of the microphone (micarrays.py n_capsules) */
  int32_t n_irs;            /* N: 0 = no IR, dry audio tiled over C channels (:572-577); 1 = static (:564-569);
                               >1 = moving, time-variant convolution (:580-587);
                               -1 = already rendered: `spatial` is an INPUT (C, n_out) that is only mixed into its
                               scene (generate_scene_audio_from_events called on cached event.spatial_audio) */
  int64_t n_ir_samples;     /* Lh */
  const int32_t* ir_frames; /* HOST pointer, (n_irs): STFT frame at which each IR starts, i.e.
                               int(np.round((ir_times*sr + 128)/128)) of synthesize.py:169; NULL unless n_irs > 1 */
  int32_t n_frames;         /* moving only: min(n_audio_frames, n_weight_frames) of synthesize.py:208-210 */
  int32_t normalize_irs;    /* 1: apply normalize_irs (synthesize.py:404-428 as called at :560); 0: use taps as given */
  int32_t gain_mode;        /* ALR_GAIN_EVENT or ALR_GAIN_NONE */
  double snr;               /* event.snr */
  double ref_db;            /* scene.ref_db */
  /* ---- dry / direct-path audio, compute_dry_audio (synthesize.py:432-504); dry == NULL disables it ---- */
  int32_t dry_channel;      /* event.ref_ir_channel */
  int32_t dry_low;          /* int(low_ms * sr / 1000) */
  int32_t dry_high;         /* int(high_ms * sr / 1000) */
  int32_t reserved0;
  float* dry;               /* out (n_audio + n_ir_samples - 1) -> event._spatial_audio_dry[mic] */
  /* ---- output ---- */
  float* spatial;           /* out (n_channels, n_out) row-major -> event.spatial_audio[mic]. With ALR_MEM_HOST it may be
                               NULL for an event that is mixed into a scene (scene >= 0): the event is rendered and mixed
                               on the device and never copied back (dataset generation keeps only the mix) */
  int64_t n_out;            /* samples per channel to produce: n_audio for render_event_audio (pad_or_truncate,
                               :590); n_audio+n_ir_samples-1 for the raw static convolution */
  /* ---- placement in a scene mix (generate_scene_audio_from_events, synthesize.py:359-378) ---- */
  int32_t scene;            /* index into the scenes array of the same call, or -1: do not mix */
  int32_t reserved1;
  int64_t scene_start;      /* max(0, round(event.scene_start*sr))                 (:361) */
  int64_t scene_end;        /* min(round(event.scene_end*sr), n_samples); events with end<=start are skipped (:362-369) */
  /* ---- optional device-side augmentation of the dry audio (f1) ---- */
  const alr_aug_op* aug_ops; /* HOST array (n_aug_ops), applied in order; NULL = none */
  int32_t n_aug_ops;
  int32_t normalize_audio;   /* 1: divide by max(|x| + tiny) afterwards, as Event.load_audio(normalize=True) */
  float* audio_out;          /* optional out (n_audio): the augmented / normalised dry audio (== event.audio) */
} alr_event;

/* One (scene, microphone) mixdown = one iteration of the mic loop of generate_scene_audio_from_events (:325-401). */
typedef struct alr_scene {
  int32_t n_channels;            /* C */
  int32_t n_ambience;            /* number of ambience layers (scene.ambience, :335-356) */
  int64_t n_samples;             /* T = round(scene.duration * sr) (:331) */
  const float* const* ambience;  /* HOST array of n_ambience pointers, each (C, T) float32 == Ambience.load_ambience().
                                    An entry may be NULL when ambience_seed is given: that layer is GENERATED on the device */
  const double* ambience_ref_db; /* HOST array (n_ambience) */
  float* mix;                    /* out (C, T) float32 -> scene.audio[mic]. With ALR_MEM_HOST it may be NULL when pcm16
                                    is given (only the PCM copy is downloaded) */
  const uint64_t* ambience_seed; /* optional HOST array (n_ambience), NULL = every layer is an input. For a layer whose
                                    ambience[k] is NULL: Gaussian noise, Ambience(noise="gaussian").load_ambience()
                                    (ambience.py:155-163, peak-normalised per channel :210-214), drawn on the device
                                    from Philox4x32-10 keyed with ambience_seed[k] ("next" row f3). The reference uses
                                    numpy's unseeded global generator there, so only the distribution is defined. */
  int16_t* pcm16;                /* optional out (T, C) interleaved 16-bit PCM: the sample data Scene.generate stores with
                                    sf.write(path, mix.T, sr) (core.py:1840-1847; libsndfile's default WAV subtype PCM_16,
                                    float -> short as lrintf(x * 32767) truncated to 16 bits, no clipping). NULL = none */
} alr_scene;

/* Per-event numbers the host layer needs back (always HOST memory). */
typedef struct alr_event_stats {
  double peak;        /* max|y| over the (C, n_out) block before gain  (apply_snr's denominator, clamp 1e-15 applied) */
  double mean_abs;    /* mean|y| over the same block */
  double gain;        /* total multiplier applied to y */
  double event_scale; /* db_to_multiplier(ref_db+snr, mean|apply_snr(y)|) == `event_scale` of synthesize.py:598 */
  int32_t nonfinite;  /* 1 if y held NaN/Inf (librosa.util.valid_audio would raise, :603) */
  int32_t dry_peak;   /* argmax used by compute_dry_audio (:482), -1 if no dry audio */
} alr_event_stats;

/* Timing / accounting of the last alr_render call (device times from CUDA events on the render stream). */
typedef struct alr_profile {
  double ms_total;          /* whole call on the device, incl. copies in ALR_MEM_HOST mode */
  double ms_ir_fft;         /* RIR partition spectra kernel */
  double ms_x_fft;          /* cross-fade + source block spectra kernel */
  double ms_cmac;           /* spectral multiply-accumulate kernel, moving events (k_cmac) */
  double ms_cmac_static;    /* spectral multiply-accumulate kernel, static events (k_cmac_static) */
  double ms_ifft;           /* inverse FFT + overlap-add + reductions kernel */
  double ms_mix;            /* gain + mixdown kernels (incl. ambience reduction) */
  double ms_other;
  double ms_fused;          /* persistent producer/consumer launch for moving events (k_mov_fused: RIR spectra + multiply-accumulate) */
  double ms_host_plan;       /* host time spent planning (overlaps GPU execution from the second chunk on) */
  int64_t kernel_launches;  /* kernels launched by the call */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int64_t workspace_bytes;
  int64_t n_chunks;
} alr_profile;

int alr_version(void);
const char* alr_last_error(void);
/* sizeof() of the ABI structs as compiled: 0 alr_event, 1 alr_scene, 2 alr_event_stats, 3 alr_profile, 4 alr_aug_op
 * (lets a binding verify its mirror of the layout; returns -1 for an unknown id) */
int alr_struct_size(int which);

/* device < 0: current device. Fails with ALR_ERR_NO_DEVICE when there is no GPU. */
int alr_create(int device, alr_context** out);
void alr_destroy(alr_context* ctx);

/* Upper bound for the spectra workspace (bytes); work is cut into chunks that fit. Default 16 GiB (device inputs; host inputs
 * use at most 512 MiB so that transfers pipeline). Only what a call needs is allocated. */
int alr_set_workspace_limit(alr_context* ctx, int64_t bytes);
/* Tuning switches (name, value); unknown names fail with ALR_ERR_INVALID.
 *   "fused"       how moving events are rendered. 0: k_ir_fft + k_ir_scale + k_cmac (RIR spectra through HBM);
 *                 1: k_mov_fused, persistent launch with FFT tasks and output-stationary multiply-accumulate tasks
 *                 exchanging the spectra through a ring (experiment: loses, profiles/r02_fused_ring.txt);
 *                 2: k_mov_sweep, warp-specialised persistent kernel (FFT producer warps, a TMA copy warp and
 *                 input-driven sweeper warps with shared-memory accumulators), spectra stay in an L2-resident ring
 *   "ring_bytes"  size of that ring (default 64 MiB)
 *   "lookahead"   runs of output blocks whose RIR spectra are produced ahead of their consumers (default 2)
 *   "small_rir"   1 (default): static renders whose effective RIR fits one partition (short RIRs with at most 2 capsules, and
 *                 the dry / direct-path sub-events of compute_dry_audio) go through k_small_rir, which keeps every spectrum
 *                 in registers; 0: the general partitioned pipeline
 *   "cmac_merge"  1: the multiply-accumulate CTAs of moving and of static events share one grid, interleaved at their ratio
 *                 (k_cmac_both; experiment, no gain: 5.89 vs 4.27 + 1.59 ms); 0 (default): two launches, k_cmac then
 *                 k_cmac_static
 *   "mix_group"   scenes per ambience-reduction + mixdown launch group, sized so that a group's ambience stays in L2
 *                 between the two passes; 0 = all scenes in one group */
int alr_set_option(alr_context* ctx, const char* name, int64_t value);
/* 1: record per-kernel CUDA-event timings into alr_profile (adds event records, no syncs). Default 0. */
int alr_set_profiling(alr_context* ctx, int enable);

/*
 * Render n_events events and mix n_scenes scenes.
 *   mem_space   ALR_MEM_DEVICE: audio / irs / ambience / spatial / dry / mix are device pointers on ctx's GPU;
 *               ALR_MEM_HOST:   they are host pointers; the call stages them through pinned memory, copies
 *                               inputs host->device and results device->host on `stream`.
 *   stats       HOST array (n_events) or NULL.
 *   stream      a cudaStream_t (as void*), NULL = the legacy default stream.
 * The call returns after the work has been enqueued AND completed (it synchronises `stream`): results and
 * stats are valid on return.
 */
int alr_render(alr_context* ctx, const alr_event* events, int64_t n_events, const alr_scene* scenes,
               int64_t n_scenes, int mem_space, alr_event_stats* stats, void* stream);

int alr_get_profile(alr_context* ctx, alr_profile* out);

/*
 * f4, a consumer of the mix: the STFT / visibility front-end of acoustic imaging — extract_visibilities + form_visibility
 * (imaging.py:455-719) for all frequency bands of get_visibility_matrix (:775-853) in one call.
 *   mix        (C, T) float32 = scene.audio[mic]  (the reference passes its transpose, (samples, channels))
 *   N = int(rate * t_sti) samples per short-time frame, n_stf = T / N frames, each multiplied with a Tukey(alpha) window
 *   (scipy.signal.windows.tukey, sym=True; alpha = 1 in form_visibility) and transformed with an N-point DFT; per band b
 *   the bins [int((fc[b] - bw/2) N / rate), int((fc[b] + bw/2) N / rate)] (Python slice semantics) are summed to one
 *   complex value per channel, S[f, b, c]; the visibility of a stationarity block of n_sti_per_block frames is
 *   V[blk, b, i, j] = sum_f conj(S[f, b, i]) S[f, b, j].
 *   out        (n_blocks, n_bands, C, C) complex128 as (re, im) float64 pairs, n_blocks = n_stf / n_sti_per_block.
 * Only the needed bins are evaluated (one modulated window per band, built in float64 on the host), accumulation is
 * float64. mem_space tells where mix and out live.
 */
int alr_visibilities(alr_context* ctx, const float* mix, int32_t n_channels, int64_t n_samples, double rate, double t_sti,
                     const double* fc, int32_t n_bands, double bw, int32_t n_sti_per_block, double tukey_alpha,
                     double* out, int mem_space, void* stream);

/* ---- unit-test hooks for the FFT core (device pointers) -------------------------------------------------
 * Forward: n_blocks real blocks of `n_valid` (<= partition) samples each, implicitly zero-padded to 2*partition,
 * through the negacyclic fold + twist transform of the kernels (csrc/alr_fft.cuh): `partition` ordinary complex
 * values per block, (re, im) interleaved; bin k is the block's polynomial evaluated at exp(i*pi*(4k+1)/(2*partition)).
 * Inverse: such spectra -> 2*partition real samples per block (first half, then the overlap tail), scaled by
 * 1/partition, so that irfft(rfft(a) * rfft(b)) is the linear convolution of two zero-padded blocks. */
int alr_partition_size(void);

/* Page-locked host memory for the buffers of ALR_MEM_HOST calls (cudaHostAlloc / cudaFreeHost). Optional: any host
 * pointer works, pinned ones are copied at the full PCIe rate and asynchronously. The reference keeps its arrays in
 * ordinary numpy memory; the Python host layer converts the float64 RIRs to float32 straight into such blocks. */
int alr_pinned_alloc(alr_context* ctx, size_t bytes, void** out);
void alr_pinned_free(alr_context* ctx, void* p);
int alr_debug_rfft(alr_context* ctx, const float* in, int64_t n_blocks, int64_t in_stride, int32_t n_valid,
                   float* spec_out, void* stream);
int alr_debug_irfft(alr_context* ctx, const float* spec_in, int64_t n_blocks, float* out, void* stream);

/* Host-only (no GPU needed): runs the planner on ONE event and returns its partition plan.
 *   header[8] = {K, B_valid, B_out, n_valid, xlimit, n_irs, n_wband, n_lrange}
 *   irs       n_irs x 6 int32: {xb0, xnb, xslot, woff, jmin, nrows} per IR
 *   wband     cross-fade weight bands (float32), lrange: {lmin, lmax} per output block.
 * Buffers that are too small (capacity in elements) make the call fail with ALR_ERR_INVALID. */
int alr_debug_plan(const alr_event* ev, int32_t* header, int32_t* irs, int64_t irs_cap, float* wband,
                   int64_t wband_cap, int32_t* lrange, int64_t lrange_cap);

/* Host-only: the task queue / production plan of the persistent moving-event launches for a list of events
 * (mode 1 = k_mov_fused, 2 = k_mov_sweep), for CPU tests of the planner's invariants.
 *   header[5]   {fused RIRs, tasks, ring slots, fused events, sweeper slots used}
 *   tasks       4 int32 per task: {type (0 P, 1 C), event, RIR | run, capsule | sub-tile}
 *   per_ir      10 int32 per fused RIR ordinal: {event, l, ring slot, slots, pop_x, pop_y, readers, ready target,
 *               production index, active source blocks}; pop_x / pop_y index production order */
int alr_debug_plan_movers(const alr_event* events, int32_t n_events, int32_t mode, int64_t ring_bytes, int32_t lookahead,
                          int32_t n_slots, int32_t* header, int32_t* tasks, int64_t tasks_cap, int32_t* per_ir,
                          int64_t per_ir_cap);

#ifdef __cplusplus
}
#endif
#endif /* ALRENDER_H */
