"""Summarise an .ncu-rep (raw page) into key per-launch metrics + top stall reasons.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--lines file.cu]   (runs here, no GPU needed)"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = raw(rep)
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        print('==', name[:140])
        for k in KEYS:
            if k in hdr:
                print('   %-70s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = [(hdr[i], float(r[i].replace(',', '') or 0)) for i in range(len(hdr))
              if 'smsp__average_warps_issue_stalled' in hdr[i] and hdr[i].endswith('_per_issue_active.ratio')]
        for k, v in sorted(st, key=lambda x: -x[1])[:7]:
            print('   stall %-40s %.2f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    if '--lines' in sys.argv:
        src_path = sys.argv[sys.argv.index('--lines') + 1]
        src = open(src_path).read().split('\n')
        out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
        secs, cur = [], None
        for r in csv.reader(io.StringIO(out)):
            if r and r[0] == 'Function Name':
                cur = {'name': r[1], 'rows': []}
                secs.append(cur)
            elif cur is not None:
                cur['rows'].append(r)
        seen = set()
        for s in secs:
            if s['name'] in seen or len(s['rows']) < 200:
                continue
            seen.add(s['name'])
            hdr = None
            per, tot = [], 0
            for r in s['rows']:
                if r and r[0] == 'Line No':
                    hdr = r
                    ie, iss = hdr.index('Instructions Executed'), hdr.index('# Samples')
                    continue
                if hdr is None or len(r) <= ie or not r[0].isdigit():
                    continue
                try:
                    n = int(r[ie]); sm = int(r[iss])
                except ValueError:
                    continue
                per.append((int(r[0]), n, sm)); tot += n
            tots = sum(x[2] for x in per) or 1
            print('=====', s['name'][:150], 'warp-instr', tot)
            for ln, n, sm in sorted(per, key=lambda x: -x[2])[:22]:
                print('%5d inst %5.1f%% samples %5.1f%%  %s' % (ln, 100 * n / max(tot, 1), 100 * sm / tots, src[ln - 1].strip()[:100] if ln <= len(src) else ''))


if __name__ == '__main__':
    main()
