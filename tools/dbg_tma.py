import sys, numpy as np, torch
sys.path.insert(0, '.')
import transtacos_retunegan_b200 as sb
from oracle import spectral_oracle as O
y = O.synth_noise(256*40-1, 3)
S, M = sb.transtacos_audio.get_specs(y)
torch.cuda.synchronize()
So, Mo = O.tt_get_specs(y)
print("rel", np.linalg.norm(S-So)/np.linalg.norm(So), np.linalg.norm(M-Mo)/np.linalg.norm(Mo))
