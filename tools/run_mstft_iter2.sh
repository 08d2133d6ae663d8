python -m pytest tests -m gpu -x -q -k "stft_loss or mstft or stft_torch or real_recording or dist or thread" 2>&1 | tail -3
bash tools/ab_mstft.sh
python tools/probe_mstft_graph.py
