cd $GRAFT_REPO_ROOT
for f in tests/test_checkpoint_gpu.py tests/test_pomdp_gpu.py tests/test_mdp_gpu.py tests/test_host_mirror.py; do
  n=$(basename $f .py)
  timeout 1500 python -X faulthandler -m pytest $f -v -m gpu --timeout=1200 > gpurun_out/dbg_$n.log 2>&1
  echo "== $f exit $?"; grep -E "PASSED|FAILED|ERROR|Fatal|Segmentation|passed|failed" gpurun_out/dbg_$n.log | tail -40
done
