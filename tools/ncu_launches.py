"""summarise an ncu --csv launch list (per-launch time, DRAM bytes, issue rate) of the solver kernels"""
import collections
import csv
import re
import sys


def num(m, n):
    v, u = m.get(n, ('0', ''))
    v = float(v.replace(',', ''))
    mult = {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3, 'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    return v * mult


def main(path, json_out=None, model=None, batch=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    data = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        key = (int(d['ID']), d['Kernel Name'][:80])
        data.setdefault(key, {})[d['Metric Name']] = (d['Metric Value'], d['Metric Unit'])
    tot = 0.0
    k3_bytes = 0.0; k3_ms = 0.0
    step = -1; step_bytes = {}; step_ms = {}; step_n = {}      # per solver step (a step starts at its linearise kernel)
    step_fl = {}                                                # executed fp64 flops of the K3 launches (dadd + dmul + 2 dfma thread instructions)
    names = {'0': 'B_first', '1': 'B', '2': 'F', '3': 'Bd', '4': 'Fd', '5': 'FDF'}
    for (i, k), m in data.items():
        kind = re.search(r'k_sweep<([\w:]+), (?:\(int\))?(\d)>', k)
        grp = 'k_ipm_group' in k or 'k_ipm_coop' in k or 'k_handover' in k or 'k_ipm_finish' in k
        name = names[kind.group(2)] if kind else ('K3group' if 'k_ipm_group' in k else 'K3coop' if 'k_ipm_coop' in k else k[:24])
        if 'k_linearize' in k:
            step += 1
        t = num(m, 'gpu__time_duration.sum'); rd = num(m, 'dram__bytes_read.sum'); wr = num(m, 'dram__bytes_write.sum')
        tot += t
        if kind or grp:
            k3_bytes += rd + wr; k3_ms += t
            if step >= 0:
                step_bytes[step] = step_bytes.get(step, 0.0) + rd + wr; step_ms[step] = step_ms.get(step, 0.0) + t
                step_n[step] = step_n.get(step, 0) + 1
                fl = (num(m, 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum') + num(m, 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum')
                      + 2.0 * num(m, 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum'))
                step_fl[step] = step_fl.get(step, 0.0) + fl
        extra = ''
        for key, lab in (('smsp__inst_executed.sum', 'inst'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'),
                         ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
                         ('smsp__thread_inst_executed_per_inst_executed.ratio', 'thr/inst'),
                         ('l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum', 'local_ld_B')):
            if key in m:
                extra += f' {lab}={num(m, key):.4g}'
        print(f'{i:3d} {name:8s} t={t:.3f}ms rd={rd / 1e9:.2f}GB wr={wr / 1e9:.2f}GB bw={(rd + wr) / max(t, 1e-9) / 1e9:.2f}TB/s{extra}')
    print(f'total {tot:.3f} ms over {len(data)} launches; K3 launches: {k3_ms:.3f} ms, {k3_bytes / 1e9:.2f} GB DRAM')
    if step_bytes:
        print('per step (K3 launches only): ' + ', '.join(f'step {q}: {step_n[q]} launches {step_ms[q]:.3f} ms {step_bytes[q] / 1e9:.2f} GB' for q in sorted(step_bytes)))
    if json_out:
        import json
        q = 0      # the first step of the capture is complete (the last one may be cut by ncu -c)
        json.dump({"model": model, "batch": batch, "dram_bytes_per_step": step_bytes[q], "k3_ms_under_ncu": step_ms[q],
                   "k3_launches_per_step": step_n[q], "source": path}, open(json_out, "w"), indent=1)
        if step_fl.get(q, 0.0) > 0.0:
            import os
            json.dump({"model": model, "batch": batch, "flops_per_step": step_fl[q], "source": path,
                       "what": "smsp__sass_thread_inst_executed_op_{dadd,dmul}_pred_on + 2 x ..._dfma_pred_on summed over the K3 launches of one step"},
                      open(os.path.join(os.path.dirname(json_out), "k3_exec_flops.json"), "w"), indent=1)


if __name__ == '__main__':
    # usage: ncu_launches.py launches.csv [out.json model batch]
    if len(sys.argv) >= 5:
        main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]))
    else:
        main(sys.argv[1])
