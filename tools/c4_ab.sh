for t in 64 128 256; do
  echo "threads=$t"
  GRIMB_THREADS=$t C4_SUBJECTS=6000 C4_SAMPLE=2 python tools/run_configs.py c4heavy 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(' c4heavy', round(d['gpu_subjects_per_s']), d['gpu_abi_seconds'], d['parity_identical_on_sample'])"
  GRIMB_THREADS=$t python tools/profile_general.py c4 20000 2>&1 | tail -2
  GRIMB_THREADS=$t GRIMB_FAST=0 python tools/profile_general.py c3 60000 2>&1 | tail -2
done
