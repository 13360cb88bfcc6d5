#!/bin/bash
# Last pass of the round: the driver-style GPU suite (pair-GEMM tests included), smoke, the bench line, and one extra
# bench line with the CTA-pair GEMMs switched on (selection rule of launch_gemm).
mkdir -p gpurun_out
echo "=== GPU suite"; timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/tests.log
echo "=== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "=== bench"; timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-300
echo "=== bench with STTS_GEMM_2CTA=1"
STTS_GEMM_2CTA=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_pair.json | cut -c1-300
echo "=== path parity with STTS_GEMM_2CTA=1"
STTS_GEMM_2CTA=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vocoder_vs_reference or config1_end_to_end or full_size or codec_encoder or ragged_batch" 2>&1 | tail -3 | tee gpurun_out/pair_parity.log
echo "=== clone path timing"
timeout 120 python - <<'PY' 2>&1 | tail -4 | tee gpurun_out/clone_timing.log
import time, numpy as np, torch
from smalltts_b200.infer import SmallTTS
t = SmallTTS.synthetic(encoder_seed=2)
wav = (0.2 * np.random.default_rng(0).standard_normal((1, 3 * 44100))).astype(np.float32)
xd = torch.from_numpy(wav).cuda()
for _ in range(3): t.engine.resample(xd, 44100, 24000)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): y = t.engine.resample(xd, 44100, 24000)
torch.cuda.synchronize(); print("resample 3 s 44.1k->24k (device buffers): %.1f us per call" % ((time.perf_counter() - t0) / 20 * 1e6))
for _ in range(2): t.clone_voice(wav, 44100)
t0 = time.perf_counter()
for _ in range(10): r = t.clone_voice(wav, 44100)
print("clone_voice (H2D + resample + encoder + D2H): %.2f ms per call, codec_enc %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3, t.engine.timings()["codec_enc_ms"]), r.shape)
PY
