# Short round-end refresh on one B200 (after tools/evidence_run.sh): GPU tests, the default bench line, launch list,
# ncu --set full of the kernels changed last (statistics, row bases, search, epilogue), the 10 M pair whole and as the
# slabs of 2 / 4 / 8 ranks (one GPU, no exchange: --shard-of).  Reports become text on the box.
cd $GRAFT_REPO_ROOT
O=gpurun_out
SO=open_pcc_metric_b200/libpccm.so
post() {
    rep=$O/$1.ncu-rep; stem=$1; shift
    [ -f $rep ] || { echo "no $rep"; return; }
    ncu -i $rep --page details > $O/${stem}_details.txt 2>/dev/null
    ncu -i $rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread > $O/${stem}_dram.csv 2>/dev/null
    for kk in "$@"; do          # kernel[:mangled template arguments]
        k=${kk%%:*}; x=""; [ "$kk" != "$k" ] && x=${kk#*:}
        ncu -i $rep --page source --csv --print-source=sass -k regex:$k > $O/_sass.csv 2>/dev/null
        python tools/ncu_by_line.py $O/_sass.csv $SO $k $x > $O/${stem}_${k}_by_line.txt 2>&1
        rm -f $O/_sass.csv
    done
    rm -f $rep
}
sum() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('dev ms',round(d['ms_per_step'],4),'frac',round(r['frac'],4), 'stage ms', round(r.get('avg_launch_ms',0),4), {k:round(v,4) for k,v in d['stage_ms_per_step'].items() if v}, d['gpu_launches'])"; }
python -m pytest tests -m gpu -x -q > $O/r2_gputests.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2_gputests.log
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/ncu_l.log 2>&1; echo "ncu list rc=$?"
K='stats_kernel|vx_rowbase_kernel|vx_search_kernel|vx_epilogue_kernel'
read SKIP CNT <<< $(python tools/step_window.py $O/r2_launches.csv "$K")
echo "profiled launches per evaluation: $CNT (after $SKIP)"
ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $SKIP --launch-count $CNT -o $O/r2_final_full -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/ncu_f2.log 2>&1; echo "ncu full rc=$?"
post r2_final_full vx_epilogue_kernel:ILb1ELb0 stats_kernel:ILi2E
S="python bench.py --config split --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
$S > $O/r2_bench_10m_n1.json 2> $O/r2_bench_10m_n1.err; echo "10m rc=$?"
{ echo "10 M + 10 M vox12 pair on ONE GPU: whole, and what rank R of W does (bench.py --config split --shard-of R,W; no exchange)";
  echo "whole"; cat $O/r2_bench_10m_n1.json | sum;
  for sh in 0,2 1,4 0,8 3,8 7,8; do echo "rank,world = $sh"; $S --shard-of $sh 2>/dev/null | sum; done; } > $O/r2_shard_emulation.txt
cat $O/r2_shard_emulation.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
du -sh $O
