#!/bin/bash
# where do k_collide's DRAM reads come from?  (a) isolated launches on the half-colliding 31.5 k set; (b) inside the GD bench step with / without an L2 persisting window
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out; O=gpurun_out/k3_dram2.txt; : > $O
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
echo "== isolated, feas31k" >> $O
timeout 600 ncu --metrics $M --clock-control none --cache-control none -k regex:"^k_collide$" -s 19 -c 2 python tools/k3_probe.py CKC_WIDE_CNT 8 2>&1 | grep -E "dram__|lts__|gpu__time|l1tex" >> $O
for P in 0 1; do
echo "== GD bench step, CKC_L2_PERSIST=$P" >> $O
CKC_L2_PERSIST=$P timeout 600 ncu --metrics $M --clock-control none --cache-control none -k regex:"^k_collide$" -s 6 -c 2 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu 2>&1 | grep -E "dram__|lts__|gpu__time|l1tex" >> $O
CKC_L2_PERSIST=$P timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('gd value %.4e ms/step %.3f share %s'%(d['value'],d['ms_per_step'],d['roofline']['stage_share_of_step']))" >> $O
done
cat $O
