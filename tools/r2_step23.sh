#!/bin/bash
# A/B capture of the marching wgrad against the N = 192 + 96 variant (FETAL_B200_WGRAD_GEN=2). That variant was removed
# after this measurement (profiles/README.md keeps the numbers); GEN=3 now selects the single-slab variant.
OUT=gpurun_out
mkdir -p $OUT; rm -f $OUT/s23_*
for g in 1 2; do
FETAL_B200_WGRAD_GEN=$g timeout 600 ncu --set full --clock-control none -k regex:"wgrad_march" --launch-skip 1 -c 1 -o $OUT/s23_g$g \
   python tools/bench_layers.py wgrad 8 dec0b > $OUT/s23_ncu_g$g.log 2>&1
ncu -i $OUT/s23_g$g.ncu-rep --page raw --csv > $OUT/s23_g${g}_raw.csv 2>/dev/null
done
rm -f $OUT/*.ncu-rep
python - <<'PY'
import csv
for g in (1,2):
    rows=list(csv.reader(open('gpurun_out/s23_g%d_raw.csv'%g)))
    hdr,r=rows[0],rows[2]
    def v(n): return r[hdr.index(n)] if n in hdr else None
    print('gen',g,'dur',v('gpu__time_duration.sum'),'pipe',v('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
      'l1tex',v('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),'lts',v('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
      'dram',v('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),'dramrd',v('dram__bytes_read.sum'))
    for h in hdr:
        if ('shared' in h and ('wavefronts' in h or 'throughput' in h or 'bank' in h)) or 'smsp__inst_executed.sum'==h or 'l1tex__data_pipe_lsu_wavefronts.sum'==h or 'lts__t_sectors_op_read.sum'==h or 'lts__t_bytes.sum'==h or 'sm__inst_executed_pipe_uniform' in h or 'l1tex__m_xbar2l1tex_read_bytes' in h:
            print('    ',h,v(h))
PY
