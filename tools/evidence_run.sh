# Round-end evidence on one B200: GPU tests, bench lines of every workload, reference arm, ncu launch list and
# --set full captures.  The .ncu-rep files are turned into text ON THE BOX and deleted (gpurun_out/ travels back
# only while it stays under 64 MiB).
cd $GRAFT_REPO_ROOT
O=gpurun_out
SO=open_pcc_metric_b200/libpccm.so
post() {    # <report stem> [kernel for a by-line view ...]
    rep=$O/$1.ncu-rep; stem=$1; shift
    [ -f $rep ] || { echo "no $rep"; return; }
    ncu -i $rep --page details > $O/${stem}_details.txt 2>/dev/null
    ncu -i $rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread > $O/${stem}_dram.csv 2>/dev/null
    for k in "$@"; do
        ncu -i $rep --page source --csv --print-source=sass -k regex:$k > $O/_sass.csv 2>/dev/null
        python tools/ncu_by_line.py $O/_sass.csv $SO $k > $O/${stem}_${k}_by_line.txt 2>&1
        rm -f $O/_sass.csv
    done
    rm -f $rep
}
python -m pytest tests -m gpu -x -q > $O/r2_gputests.log 2>&1; echo "pytest rc=$?"
python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/ncu_l.log 2>&1; echo "ncu list rc=$?"
K='vx_|stats_|finalize|pack_rgb|zero_words|gather'
read SKIP CNT <<< $(python tools/step_window.py $O/r2_launches.csv "$K")
echo "launches per evaluation: $CNT (profiling after $SKIP)"
ncu --set full --clock-control none --import-source on -k regex:"$K" --launch-skip $SKIP --launch-count $CNT -o $O/r2_step_full -f python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $O/ncu_f.log 2>&1; echo "ncu full rc=$?"
post r2_step_full vx_search_kernel vx_epilogue_kernel vx_place_kernel
python bench.py --config 3 --steps 5 --warmup 3 > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err; echo "c3 rc=$?"
python bench.py --config 4 --steps 3 --warmup 3 > $O/r2_bench_c4.json 2> $O/r2_bench_c4.err; echo "c4 rc=$?"
python bench.py --config split --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_10m_n1.json 2> $O/r2_bench_10m_n1.err; echo "10m rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'normals_int|knn_self|vx_selfnn' -c 4 -o $O/r2_c3_full -f python bench.py --config 3 --points 1000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
post r2_c3_full normals_int_kernel
ncu --set full --clock-control none --import-source on -k regex:'pair_query|knn_self' -c 3 -o $O/r2_c5_full -f python bench.py --config 5 --points 2000000 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
post r2_c5_full pair_query_kernel
timeout 600 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_c5.json 2> $O/r2_bench_c5.err; echo "c5 rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
du -sh $O
