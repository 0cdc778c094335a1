# final check of a build on one B200: GPU test suite, smoke(), the default bench line and the reference arm
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > $O/r02z_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02z_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -12 $O/r02z_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r02z_b1.json 2> $O/r02z_b1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02z_ref.json 2> $O/r02z_ref.err; echo "ref rc=$?"
tail -c 600 $O/r02z_b1.err
