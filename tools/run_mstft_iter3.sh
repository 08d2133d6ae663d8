python -m pytest tests -m gpu -x -q -k "stft_loss or mstft or stft_torch or real_recording" 2>&1 | tail -2
for e in "" SB200_MSTFT_RESIDENT=0 SB200_MSTFT_SINGLE=1; do echo "== ${e:-default}"; env $e python tools/probe_mstft_graph.py; done
