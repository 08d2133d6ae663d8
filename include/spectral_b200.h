/* spectral_b200.h -- C ABI of the B200-native spectral front/back end.
 *
 * Drop-in boundary for the spectral hot path of Kahsolt/TransTacoS-RetuneGAN.  The reference has no
 * FFI layer: its boundary is the module-level Python API of transtacos/audio.py,
 * retunegan/audio.py and retunegan/models/loss.py.  The Python mirror of that API lives in
 * transtacos-retunegan_b200/ and calls ONLY the entry points below (ctypes, raw device pointers,
 * cudaStream_t last).  Every entry point cites the reference call it replaces.
 *
 * Conventions
 *   - return value: 0 = ok, negative = sb200_status; sb200_last_error_string() describes the failure.
 *   - all data pointers are DEVICE pointers unless the name ends in _host.
 *   - the caller allocates every input, output and workspace; the library owns only the immutable
 *     plan tables (window, twiddles, banded mel filterbank) behind sb200_plan.
 *   - spectrogram-shaped arrays are FRAME-MAJOR: element (frame t, bin k) at [t*F + k], F = n_fft/2+1,
 *     i.e. the memory order librosa.stft(order='F') / torch.stft produce; logically [F, T] with strides (1, F).
 *   - launches are asynchronous on `stream`; errors of the launch itself are returned, execution errors
 *     surface at the caller's next synchronisation (as in torch).
 *   - a batch is described by sb200_batch: either uniform (B rows of the same length) or ragged
 *     (offset tables on the device).  Utterances never interact.
 */
#ifndef SPECTRAL_B200_H_
#define SPECTRAL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb200_plan sb200_plan;
typedef void* sb200_stream; /* cudaStream_t */

typedef enum {
  SB200_OK = 0,
  SB200_ERR_INVALID = -1,      /* bad argument / unsupported configuration (Python: ValueError) */
  SB200_ERR_CUDA = -2,         /* CUDA runtime error (Python: RuntimeError) */
  SB200_ERR_UNSUPPORTED = -3,  /* valid in the reference but not built here (Python: NotImplementedError) */
} sb200_status;

typedef enum { SB200_WIN_HANN = 0, SB200_WIN_HAMMING = 1, SB200_WIN_BLACKMAN = 2, SB200_WIN_BARTLETT = 3 } sb200_window;

/* Mirrors the hparam.py fields the path reads (transtacos/hparam.py:5-17 == retunegan/hparam.py:3-15,36-37). */
typedef struct {
  int32_t sample_rate; /* 22050 */
  int32_t n_fft;       /* 2048 | 1024 | 512 */
  int32_t win_length;  /* must be n_fft/2 (all reference configurations) */
  int32_t hop_length;  /* even, 1 <= hop <= win_length */
  int32_t n_mel;       /* <= 128 */
  float fmin, fmax;    /* fmax < sample_rate/2 (transtacos/audio.py:160) */
  int32_t mel_htk;     /* 0 = Slaney scale (reference default), 1 = HTK (retunegan/hparam.py:37) */
  int32_t window;      /* sb200_window; scipy.signal.get_window(name, win, fftbins=True) */
} sb200_config;

/* Batch of independent utterances.  Uniform: off == NULL, every row has `len` samples at stride
 * `stride` and n_frames = 1 + len/hop.  Ragged: device tables (int64) sig_off[B] (first sample of row b),
 * sig_len[B], frame_off[B+1] (prefix sum of frames; row b owns output frames [frame_off[b], frame_off[b+1])),
 * item_off[B+1] (prefix sum of ceil(frames_b / frames_per_pass), see sb200_plan_frames_per_pass).
 * The sample buffer need not be aligned; when x is 16-byte aligned the hot feature kernel stages samples with bulk-tensor (TMA) copies. */
typedef struct {
  int32_t B;
  int64_t len, stride;       /* uniform */
  const int64_t* sig_off;    /* ragged (device) or NULL */
  const int64_t* sig_len;
  const int64_t* frame_off;
  const int64_t* item_off;
  int64_t total_frames;      /* ragged: frame_off[B]; uniform: ignored */
  int64_t total_items;       /* ragged: item_off[B];  uniform: ignored */
  int64_t total_samples;     /* ragged: samples the flat buffer holds from x (>= sig_off[b] + sig_len[b] for every b), or 0 =
                                unknown (then the feature kernel gathers with plain loads instead of bulk-tensor copies); uniform: ignored */
} sb200_batch;

/* Output transform of a magnitude-like value v:  raw v, or a*log2(max(floor, v)) + b.
 *   get_specs  (transtacos/audio.py:73-77,177-193):  a = 2*max_abs*20*log10(2)/(-min_db), b = 2*max_abs*(-ref-min_db)/(-min_db)-max_abs, floor = 1e-5
 *   get_mag/get_mel (retunegan/audio.py:116-128):    a = ln 2, b = 0, floor = 1e-5 (clamp_low) or 0 */
typedef struct {
  int32_t log;  /* 0 = raw */
  float a, b, floor;
} sb200_scale;

/* ---- library ------------------------------------------------------------------------------------ */
const char* sb200_version(void);
const char* sb200_last_error_string(void);       /* thread-local */
/* Number of kernel launches issued by this library in this process (bench.py "gpu_launches"). */
int64_t sb200_launch_count(void);

/* ---- plans ------------------------------------------------------------------------------------------
 * Replaces the import-time / lazily cached tables of the reference: mel_basis (retunegan/audio.py:20,
 * transtacos/audio.py:157-162), window_fn_torch / mel_basis_torch (retunegan/audio.py:25-26,153-159).
 * Created on the current CUDA device. */
int sb200_plan_create(const sb200_config* cfg, sb200_plan** out);
int sb200_plan_destroy(sb200_plan* plan);
int sb200_plan_frames_per_pass(const sb200_plan* plan);             /* frames per work item: 4096 / n_fft */
/* Dense float32 filterbank [n_mel, F] == librosa.filters.mel(sr, n_fft, n_mel, fmin, fmax) (host buffer). */
int sb200_plan_mel_basis_host(const sb200_plan* plan, float* out_host);
/* float32 window [win_length] (host buffer). */
int sb200_plan_window_host(const sb200_plan* plan, float* out_host);

/* ---- STFT magnitude + mel (+ optional complex STFT) -----------------------------------------------
 * One fused launch: [pre-emphasis ->] reflect pad -> frame gather -> window -> real FFT -> |.| ->
 * banded mel projection -> output transforms.
 *   transtacos/audio.py:73-77  get_specs   (preemph = 0.97, mag/mel = dB-normalise)
 *   retunegan/audio.py:116-128 get_mag / get_mel (preemph = 0, ln + clip)
 *   retunegan/audio.py:161-168 torch.stft + abs (raw)             [complex `spec` output = D]
 * x: float32 samples.  mag [frames, F], mel [frames, n_mel], spec [frames, F] (float2 re,im); any of the
 * three may be NULL.  preemph = 0 disables the FIR x[n] - k x[n-1] (zero initial state, applied BEFORE
 * reflect padding as scipy.signal.lfilter does at transtacos/audio.py:64-66). */
int sb200_stft_features(const sb200_plan* plan, const float* x, const sb200_batch* batch, float preemph,
                        sb200_scale mag_scale, sb200_scale mel_scale, float* mag, float* mel, float* spec,
                        sb200_stream stream);

/* Banded mel projection of an arbitrary frame-major array: out[t, m] = sum_k basis[m, k] * in[t, k]
 * (retunegan/audio.py:21 mag_to_mel = np.dot(mel_basis, x); transtacos/audio.py:154-155). */
int sb200_mel_project(const sb200_plan* plan, const float* in, int64_t frames, sb200_scale scale, float* out,
                      sb200_stream stream);

/* _mel_to_linear (transtacos/audio.py:164-175): out[t, :] = linear_basis @ mel[t, :] with
 * linear_basis = mel_basis^T diag(1 / colsum(mel_basis mel_basis^T)); the first step of inv_mel (:100-104).
 * mel [frames, n_mel] -> out [frames, F], frame-major. */
int sb200_mel_to_linear(const sb200_plan* plan, const float* mel, int64_t frames, float* out, sb200_stream stream);

/* Element-wise helpers of the inverse path, fused into one launch each:
 *   mode 0: out = 10^((in + max_abs)*(-min_db)/(2*max_abs) + min_db + ref) / 20) ^ power
 *           (transtacos/audio.py:80-82 spec_to_natural_scale, then S ** gl_power at :96)
 *   mode 1: out = exp(in) ^ power      (retunegan/audio.py:140 np.exp(mag), :132 S ** gl_power)
 *   mode 2: out = in ^ power           (S ** gl_power after fix_zero_DC, transtacos/audio.py:95-96)
 * p0..p2 = (max_abs, min_db, ref_db) for mode 0. */
int sb200_spec_to_amplitude(const float* in, int64_t n, int32_t mode, float p0, float p1, float p2, float power,
                            float* out, sb200_stream stream);

/* ---- frame statistics sharing the STFT framing ------------------------------------------------------
 * rms[t]: librosa.feature.rms(y, frame_length, hop_length) (center=True, reflect padding) -- get_c0 at
 *   transtacos/audio.py:112-114 and retunegan/audio.py:103-105; with frame_length 512 / hop 128 it is the energy
 *   track of librosa.effects.trim (trim_silence, transtacos/audio.py:59-61).
 * zcr[t]: librosa.feature.zero_crossing_rate(y, frame_length, hop_length) (center=True, edge padding, threshold 1e-10)
 *   -- get_zcr at retunegan/audio.py:98-100.
 * The batch describes the signals; a row of length len has 1 + len / hop_length frames.  For a ragged batch
 * frame_off must hold the frame offsets FOR THIS hop_length.  Either output may be NULL. */
int sb200_frame_stats(const float* x, const sb200_batch* batch, int32_t frame_length, int32_t hop_length, float* rms,
                      float* zcr, sb200_stream stream);

/* librosa.effects.trim bounds (trim_silence, transtacos/audio.py:59-61) from the RMS track sb200_frame_stats wrote with
 *   frame_length 512 / hop 128: frames whose power is within top_db of the row's loudest frame are non-silent
 *   (power_to_db(ref = max, amin = 1e-10, top_db = None) > -top_db).  frame_off: device [B+1] prefix sums of the rows' frame
 *   counts, or NULL with frames_per_row for a uniform batch.  bounds (device, int64 [B, 2]): first non-silent frame and last
 *   non-silent frame + 1 of every row, (0, 0) for an all-silent row; the caller turns frames into samples (x hop, clipped to
 *   the length). */
int sb200_trim_bounds(const float* rms, const int64_t* frame_off, int64_t frames_per_row, int32_t B, float top_db,
                      int64_t* bounds, sb200_stream stream);

/* f0[t]: librosa.yin(y, fmin, fmax, sr, frame_length, hop_length=hop_length) (win_length = frame_length / 2, trough
 *   threshold 0.1, center=True, reflect padding) -- get_f0 at transtacos/audio.py:107-109.  Same batch convention as
 *   sb200_frame_stats.  frame_length <= 4096. */
int sb200_yin(const float* x, const sb200_batch* batch, int32_t sample_rate, float fmin, float fmax, int32_t frame_length,
              int32_t hop_length, float trough_threshold, float* f0, sb200_stream stream);

/* ---- waveform max-pool losses (retunegan/models/loss.py:66-82) ---------------------------------------
 * mode 0: envelope_loss = mean|MaxPool(y) - MaxPool(y_g)| + mean|MaxPool(-y) - MaxPool(-y_g)|
 * mode 1: dynamic_loss  = mean||MaxPool(y) + MaxPool(-y)| - |MaxPool(y_g) + MaxPool(-y_g)||
 * MaxPool = nn.MaxPool1d(pool_k) (stride pool_k, no padding; retunegan/hparam.py:90 envelope_pool_k = 160).
 * y, y_g [B, T]; loss: device scalar; grad_yg [B, T] (may be NULL): d loss / d y_g for a unit upstream gradient.
 * workspace: sb200_pool_loss_workspace_bytes() bytes. */
int64_t sb200_pool_loss_workspace_bytes(void);
int sb200_pool_loss(const float* y, const float* y_g, int32_t B, int64_t T, int32_t pool_k, int32_t mode, float* loss,
                    float* grad_yg, void* workspace, sb200_stream stream);

/* ---- pre-emphasis filters (transtacos/audio.py:64-70) ---------------------------------------------
 * preemphasis: y[n] = x[n] - k x[n-1];  inv_preemphasis: y[n] = x[n] + k y[n-1] (parallel scan). */
int sb200_preemphasis(const float* x, const sb200_batch* batch, float k, float* y, sb200_stream stream);
int sb200_inv_preemphasis(const float* x, const sb200_batch* batch, float k, float* y, sb200_stream stream);

/* ---- ISTFT / Griffin-Lim ----------------------------------------------------------------------------
 * Signal batches here are described by frames: row b has n_frames_b = frame_off[b+1]-frame_off[b] (uniform:
 * `len` holds n_frames) and produces out_len_b samples: hop*(n_frames_b-1) if length <= 0
 * (librosa.istft length=None) else `length` (uniform) / sig_len[b] (ragged).  Output row b starts at
 * b*stride (uniform) or sig_off[b] (ragged).
 *
 * sb200_griffinlim_workspace_bytes: bytes of `workspace` needed for a batch of n_rows utterances with total_frames frames
 *   in all (four signal buffers of total_frames*hop + n_rows*win floats; form 1 adds the previous spectrum).
 * sb200_istft: y = librosa.istft(spec) (transtacos/audio.py:147-148), spec complex [frames, F].
 * sb200_griffinlim:
 *   form 0 (transtacos/audio.py:130-140 _griffin_lim): angles = exp(i angle(STFT(y))), no momentum;
 *   form 1 (retunegan/audio.py:131-136 librosa.griffinlim): angles = c/(|c|+1e-16), c = rebuilt - momentum/(1+momentum)*tprev.
 *   S: magnitudes [frames, F] (already raised to gl_power); init_phase: [frames, F] values u in [0,1)
 *   (the np.random.rand draw of the reference; initial angles = exp(2 pi i u)).  n_iter iterations, then
 *   the final ISTFT.  inv_preemph != 0 additionally applies inv_preemphasis to the result
 *   (transtacos/audio.py:96).  y: float32 output. */
int64_t sb200_griffinlim_workspace_bytes(const sb200_plan* plan, int64_t total_frames, int32_t n_rows, int32_t form);
int sb200_istft(const sb200_plan* plan, const float* spec, const sb200_batch* frames_batch, int64_t length,
                float* y, void* workspace, sb200_stream stream);
int sb200_griffinlim(const sb200_plan* plan, const float* S, const float* init_phase,
                     const sb200_batch* frames_batch, int64_t length, int32_t n_iter, float momentum,
                     int32_t form, float inv_preemph, float* y, void* workspace, sb200_stream stream);

/* ---- multi-resolution STFT loss (retunegan/models/loss.py:22-62, retunegan/audio.py:150-170) -------
 * plans[n_res]: one plan per (n_fft, win, hop) of hp.multi_stft_params.  y, y_g: [B, T] float32.
 * Forward: loss (device scalar, may be NULL) = 1/n_res * sum_res ( mean|M - M_g| + mean|ln M - ln M_g| ),
 * M = mel_basis @ |D + 1e-9|.  specs_r[res] / specs_g[res] (may be NULL): [B, 2, T'_res, F_res] frame-major
 * stacks (channel 0 = ln|D + 1e-9|, channel 1 = angle(D)/PI with PI = 3.14159265358979, retunegan/utils.py:12).
 * phd_phase != 0 selects hp.phd_input == 'phase' (generated stack reuses the real ln-magnitude, loss.py:45-47).
 * saved: caller-allocated buffer of sb200_mstft_saved_bytes() that forward fills and backward reads.
 * Backward: g_yg [B, T] = g_loss_scalar * dloss/dy_g + sum_res <g_specs_g[res], dspecs_g[res]/dy_g>
 * (g_specs_g[res] may be NULL = zero upstream).  g_loss is a device scalar pointer.  workspace: caller buffer of
 * sb200_mstft_workspace_bytes(). */
int64_t sb200_mstft_saved_bytes(const sb200_plan* const* plans, int32_t n_res, int32_t B, int64_t T);
int64_t sb200_mstft_workspace_bytes(const sb200_plan* const* plans, int32_t n_res, int32_t B, int64_t T);
int sb200_mstft_forward(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                        int64_t T, int32_t phd_phase, float* loss, float* const* specs_r, float* const* specs_g,
                        void* saved, void* workspace, sb200_stream stream);
int sb200_mstft_backward(const sb200_plan* const* plans, int32_t n_res, const float* y_g, int32_t B, int64_t T,
                         int32_t phd_phase, const float* g_loss, const float* const* g_specs_g, const void* saved,
                         float* g_yg, void* workspace, sb200_stream stream);

/* ---- get_stft_torch (retunegan/audio.py:150-170), differentiable ---------------------------------------
 * Forward, one launch: D = torch.stft(y, n_fft, hop, win, hann, center, reflect); S = |D + 1e-9|, P = angle(D) [B, T', F] and
 * M = mel_basis S [B, T', n_mel], frame-major (the [B, F, T'] / [B, n_mel, T'] tensors of the reference are their transposed
 * views); any output may be NULL.  y [B, T].
 * Backward (what torch.autograd derives from audio.py:161-168): g_y [B, T] from the upstream gradients g_S, g_P [B, T', F] and
 * g_M [B, T', n_mel] (any may be NULL = zero): gS = g_S + mel_basis^T g_M; gD = gS (D + 1e-9)/S + g_P i D/|D|^2 (0 where
 * the magnitude is 0, as torch.abs / torch.angle do); adjoint of the one-sided windowed rFFT; overlap-add with the reflect
 * padding folded back.  The analysis of y is recomputed (nothing is saved by forward).  workspace: sb200_stft_smp_workspace_bytes(). */
int64_t sb200_stft_smp_workspace_bytes(const sb200_plan* plan, int32_t B, int64_t T);
int sb200_stft_smp_forward(const sb200_plan* plan, const float* y, int32_t B, int64_t T, float* S, float* M, float* P,
                           sb200_stream stream);
int sb200_stft_smp_backward(const sb200_plan* plan, const float* y, int32_t B, int64_t T, const float* g_S, const float* g_M,
                            const float* g_P, float* g_y, void* workspace, sb200_stream stream);

/* Loss value and d loss / d y_g for a unit upstream gradient in ONE pass (retunegan/train.py:165 multi_stft_loss(...,
 * ret_loss=True) followed by :192 backward, loss-only): one launch for all resolutions, no second analysis in backward.
 * loss: device scalar; grad_yg: [B, T]; workspace: sb200_mstft_workspace_bytes(). */
int sb200_mstft_loss_and_grad(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                              int64_t T, float* loss, float* grad_yg, void* workspace, sb200_stream stream);

/* ---- DDP: the loss averaged over the ranks of one box INSIDE the reducing kernel --------------------------------------------
 * Under DDP every rank computes l_mstft on its own segments (retunegan/train.py:165 inside the DDP-wrapped step); what is logged
 * is its mean over the ranks.  Instead of an NCCL all-reduce of one scalar after the step (a host call and a launch on a step
 * that is itself bound by the host), the block that reduces the rank's loss exchanges it with the other ranks over NVLink peer
 * memory: it stores the value into every rank's exchange buffer, publishes it with a release store of an epoch counter, waits
 * for the other ranks' flags and adds the values in rank order (same result on every rank).  No launch, no host call, CUDA-graph
 * safe (the epoch lives in device memory).  A wait longer than 2 s yields NaN instead of hanging.
 *
 * Every rank creates one exchange buffer (sb200_peer_buffer_create: cudaMalloc + zero fill + CUDA IPC handle, 64 bytes), the
 * handles are exchanged by the caller (e.g. torch.distributed.all_gather_object) and opened with sb200_peer_buffer_open.
 * peer[q] = exchange buffer of rank q as a device pointer valid in THIS process (the own buffer at [rank]); all ranks must make
 * the same sequence of *_ddp calls.  world <= SB200_MAX_PEERS (one NVSwitch box); larger jobs reduce with NCCL instead. */
#define SB200_MAX_PEERS 8
typedef struct {
  void* peer[SB200_MAX_PEERS];
  int32_t rank, world;
  float* loss_global; /* device scalar out: mean of the ranks' losses */
} sb200_peer_reduce;
int64_t sb200_peer_buffer_bytes(void);
int sb200_peer_buffer_create(void** buf, void* ipc_handle_out_64_bytes /* may be NULL: single-process use */);
int sb200_peer_buffer_open(const void* ipc_handle_64_bytes, void** buf);
int sb200_peer_buffer_close(void* buf);   /* a buffer obtained from sb200_peer_buffer_open */
int sb200_peer_buffer_destroy(void* buf); /* a buffer obtained from sb200_peer_buffer_create */
/* sb200_mstft_forward / sb200_mstft_loss_and_grad (loss required) + the reduction described above; `loss` still receives the
 * rank-local value (the one the gradient belongs to), peers->loss_global the mean over the ranks. */
int sb200_mstft_forward_ddp(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                            int64_t T, int32_t phd_phase, float* loss, float* const* specs_r, float* const* specs_g,
                            void* saved, void* workspace, const sb200_peer_reduce* peers, sb200_stream stream);
int sb200_mstft_loss_and_grad_ddp(const sb200_plan* const* plans, int32_t n_res, const float* y, const float* y_g, int32_t B,
                                  int64_t T, float* loss, float* grad_yg, void* workspace, const sb200_peer_reduce* peers,
                                  sb200_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRAL_B200_H_ */
