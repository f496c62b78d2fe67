# Round-2 profile session (one B200): launch list of the bench command + ncu --set full captures of the kernels the bench
# line cites; every capture is summarised to text on the box (the .ncu-rep files exceed what travels back) and deleted.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu > $O/r02_prof_bench_plain.json 2> $O/r02_prof_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_bench_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/r02_prof_ncu_bench.log 2>&1
summ() {   # name: raw-page summary, SASS-level source page for single-kernel captures
    python scripts/ncu_summary.py /tmp/$1.ncu-rep > $O/$1.summary.txt 2>&1
    if [ "$2" = "src" ]; then ncu -i /tmp/$1.ncu-rep --page source --csv --print-source sass > $O/$1.sass.csv 2>/dev/null; fi
    rm -f /tmp/$1.ncu-rep
}
python scripts/prof_far.py cfg2 1 > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:k2_line_sum -c 1 -o /tmp/r02_k2_exact_cfg2 python scripts/prof_far.py cfg2 1 > $O/r02_prof1.log 2>&1; summ r02_k2_exact_cfg2 src
ncu --set full --import-source on --clock-control none -k regex:k2_line_sum_far -c 1 -o /tmp/r02_k2_far_cfg2 python scripts/prof_far.py cfg2 2 > $O/r02_prof2.log 2>&1; summ r02_k2_far_cfg2 src
ncu --set full --import-source on --clock-control none -k regex:k2_line_sum_far -c 1 -o /tmp/r02_k2_far_cfg5 python scripts/prof_far.py cfg5 2 > $O/r02_prof3.log 2>&1; summ r02_k2_far_cfg5 src
python scripts/prof_atm.py 2 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:"k1_prepass|k2_|k3_fold" -c 12 -o /tmp/r02_atm_far python scripts/prof_atm.py 2 > $O/r02_prof4.log 2>&1; summ r02_atm_far
ncu --set full --clock-control none -k regex:"k2_line_sum" -c 4 -o /tmp/r02_atm_exact python scripts/prof_atm.py 1 > $O/r02_prof5.log 2>&1; summ r02_atm_exact
ncu --set full --import-source on --clock-control none -k regex:k1_prepass -c 1 -o /tmp/r02_k1_cfg4 python scripts/prof_atm.py 1 > $O/r02_prof6.log 2>&1; summ r02_k1_cfg4 src
du -sh $O
