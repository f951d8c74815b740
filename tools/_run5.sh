cd $GRAFT_REPO_ROOT
for v in 1 0; do
echo "== PCCM_VOX=$v"
PCCM_VOX=$v python tools/run_configs.py c3 10000000 2>&1 | grep "iter" | cut -c1-420
done
