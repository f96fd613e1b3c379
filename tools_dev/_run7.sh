mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -s -k "variants or batched_notes or note_synthesis or rejects" > gpurun_out/t_var.log 2>&1; echo "var rc=$?"; tail -30 gpurun_out/t_var.log
