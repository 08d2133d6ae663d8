"""Achieved parity numbers quoted in DESIGN.md section 1 (GPU vs the float64 oracle): the benchmarked 64 x 5 s feature launch,
speech-like and noise inputs, RetuneGAN ln features, Griffin-Lim, the mstft loss."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
from oracle import spectral_oracle as O
ta, ra = sb.transtacos_audio, sb.retunegan_audio
L = 431 * 256 - 1
def report(tag, y):
    S, M = ta.get_specs(y, out_dtype=np.float32); So, Mo = O.tt_get_specs(y)
    for name, a, b in (("mag", S, So), ("mel", M, Mo)):
        na, nb = O.tt_spec_to_natural_scale(a.astype(np.float64)), O.tt_spec_to_natural_scale(b)
        strong = nb >= 1e-3 * nb.max()
        print(f"{tag} {name}: magnitude rel-Frobenius {np.linalg.norm(na-nb)/np.linalg.norm(nb):.2e}, max-abs/max {np.abs(na-nb).max()/nb.max():.2e}; "
              f"normalised-dB max-abs: bins within 60 dB of the peak {np.abs(a-b)[strong].max():.2e}, all bins {np.abs(a-b).max():.2e} "
              f"(weakest bin {20*np.log10(nb.min()/nb.max()):.0f} dB)")
report("speechlike 5 s", O.synth_speechlike(L, 114514))
report("noise 5 s", O.synth_noise(L, 114515))
y = O.synth_speechlike(L, 114514)
mag, mel = ra.get_mag_mel(y)
print("rtg ln-mag: amplitude rel-Frobenius", np.linalg.norm(np.exp(mag)-np.exp(O.rtg_get_mag(y)))/np.linalg.norm(np.exp(O.rtg_get_mag(y))),
      "ln-mel max-abs", np.abs(mel-O.rtg_get_mel(y)).max())
S = np.abs(O.stft(y, 2048, 256, 1024)).astype(np.float32)
u = np.random.RandomState(114514).rand(1025, 431)
for form in ("tt", "rtg"):
    if form == "tt":
        ref = O.tt_griffin_lim(S.astype(np.float64) ** 1.2, init_phase=u); out = ta._griffin_lim(S.astype(np.float64) ** 1.2, init_phase=u)
    else:
        ref = O.rtg_griffinlim(S, wavlen=L, init_angles=np.exp(2j*np.pi*u)); out = ra._griffinlim(S, wavlen=L)
    T = S.astype(np.float64) ** 1.2
    print(f"griffin-lim {form}: waveform rel-L2 {np.linalg.norm(out-ref)/np.linalg.norm(ref):.2e}, spectral convergence gpu {O.spectral_convergence(T, out):.6f} oracle {O.spectral_convergence(T, ref):.6f}")
