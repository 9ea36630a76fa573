python -m pytest tests/test_gpu_predict.py tests/test_gpu_robustness.py tests/test_legacy_art.py -m gpu -x -q 2>&1 | tail -5
echo "--- stacked (default)"; python tools/predict_probe.py 1024 4000 2>&1 | tail -4
echo "--- dmma"; AGPN_PREDICT=dmma python tools/predict_probe.py 1024 4000 2>&1 | tail -4
echo "--- C5 stacked"; python tools/predict_probe.py 4096 100000 2>&1 | tail -3
echo "--- C5 dmma"; AGPN_PREDICT=dmma python tools/predict_probe.py 4096 100000 2>&1 | tail -3
