set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02f_b1.json 2> $O/r02f_b1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02f_ref.json 2> $O/r02f_ref.err; echo "ref rc=$?"
for t in slcp lotka_volterra sir jrnnm pf_slcp; do for cm in GAUSS JAC; do
  if [ "$t$cm" = "pf_slcpJAC" ]; then continue; fi
  timeout 200 python bench.py --task $t --cov-mode $cm --steps 3 --warmup 3 --no-sweep --no-obs-sharded --no-cpu-baseline 2>/dev/null | tail -1 >> $O/r02f_tasks.jsonl
done; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f_launches_gauss.csv python bench.py --steps 2 --warmup 1 --no-sweep --no-obs-sharded --no-cpu-baseline > $O/r02f_ncu_l.log 2>&1; echo "ncu list gauss rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02f_launches_jac.csv python bench.py --cov-mode JAC --steps 2 --warmup 1 --no-sweep --no-obs-sharded --no-cpu-baseline > $O/r02f_ncu_lj.log 2>&1; echo "ncu list jac rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:traj_tc_kernel -c 1 -f -o $O/r02f_tc python bench.py --steps 1 --warmup 1 --no-sweep --no-obs-sharded --no-cpu-baseline > $O/r02f_ncu_f.log 2>&1; echo "ncu full tc rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jac_tc_kernel -c 1 -f -o $O/r02f_jac python bench.py --cov-mode JAC --steps 1 --warmup 1 --no-sweep --no-obs-sharded --no-cpu-baseline > $O/r02f_ncu_fj.log 2>&1; echo "ncu full jac rc=$?"
ls -la $O | tail -12
