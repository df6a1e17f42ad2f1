for v in 0 1 2 3 32 33 35; do echo "dbg=$v"; DPGO_SYM_DEBUG=$v timeout 100 python scripts/phase_times.py --steps 18 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step_events'],3), 'dense', round(d['us_per_dense_apply'],1), 'pz', round(d['us_per_partial_sum'],1), 'applies', d['precond_applies'])"; done
