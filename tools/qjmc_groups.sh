timeout 200 python -m pytest tests/test_gpu_z_projsum.py -q -x -k "batching or ensemble" 2>&1 | tail -2
for G in 1 2 4; do TN_QJMC_GROUPS=$G timeout 200 python tools/bench_qjmc.py --sites 64 --chi 256 --traj 64 --steps 1 --workers 64 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('groups', $G, 'traj 64:', round(d['traj_steps_per_s'],3))"; done
TN_QJMC_GROUPS=2 timeout 200 python tools/bench_qjmc.py --sites 64 --chi 256 --traj 32 --steps 1 --workers 32 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('groups 2 traj 32:', round(d['traj_steps_per_s'],3))"
bash tools/profile_r02c.sh | tail -6 | cut -c1-260
