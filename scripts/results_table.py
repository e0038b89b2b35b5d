#!/usr/bin/env python
"""DESIGN.md section 9 table from the bench lines kept under profiles/ (r2_bench_{1,2,8}gpu.json)."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load(n): return json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_%dgpu.json" % n)).read().strip().splitlines()[-1])
d1, d2, d8 = load(1), load(2), load(8)
L = ['| config | 1 GPU | 2 GPUs | 8 GPUs |', '|---|---|---|---|']
L.append('| [1] tail p = 3, descriptors/s (µs per step, fraction of HBM peak) | %.0f k (%.1f µs, %.3f) | %.2f M (%.1f µs) | %.2f M (%.1f µs, %.2f× of 1 GPU) |' % (d1['value']/1e3, 1e3*d1['ms_per_step'], d1['roofline']['frac'], d2['value']/1e6, 1e3*d2['ms_per_step'], d8['value']/1e6, 1e3*d8['ms_per_step'], d8['value']/d1['value']))
L.append('| [1] tail p = 2.7 (µs per step, fraction) | %.1f µs, %.3f | %.1f µs | %.1f µs |' % (1e3*d1['tail_p27']['ms_per_step'], d1['tail_p27']['roofline']['frac'], 1e3*d2['tail_p27']['ms_per_step'], 1e3*d8['tail_p27']['ms_per_step']))
L.append('| [1] tail end to end from host maps (PCIe-bound) | %.1f k desc/s | %.1f k | %.1f k |' % (d1['e2e']['value']/1e3, d2['e2e']['value']/1e3, d8['e2e']['value']/1e3))
for key, name in (('q10k', '[3] 10k × 1M top-100, i.i.d. rows'), ('q10k_clustered', '[3] same, clustered rows'), ('q10k_cluster_sorted', '[3] same, clustered rows stored centre by centre')):
    L.append('| %s: ms (fraction of sustained bf16 peak per GPU) | %.2f (%.2f) | %.2f (%.2f) | %.2f (%.2f) |' % ((name,) + sum(((d['search'][key]['ms_per_search'], d['search'][key]['roofline']['frac']) for d in (d1, d2, d8)), ())))
L.append('| [3] 10k queries end to end from host memory, q/s | %.0f k | %.0f k | %.2f M |' % (d1['search']['q10k']['e2e']['value']/1e3, d2['search']['q10k']['e2e']['value']/1e3, d8['search']['q10k']['e2e']['value']/1e6))
L.append('| [3] 70 × 1M top-100: ms (fraction of HBM peak per GPU) | %.3f (%.2f) | %.3f (%.2f) | %.3f (%.2f) |' % sum(((d['search']['q70']['ms_per_search'], d['search']['q70']['roofline']['frac']) for d in (d1, d2, d8)), ()))
L.append('| [2] mining 2,000 × 20,000, nnum 5 (replicas) | %.2f ms; all 2,000 sets == CPU restatement: %s; CPU %.0f q/s | %.2f ms | %.2f ms |' % (d1['mining']['ms'], d1['mining'].get('sets_identical_to_cpu_port'), d1['mining']['cpu_baseline']['value'], d2['mining']['ms'], d8['mining']['ms']))
L.append('| [4] αQE k=10 α=3 of 10k queries + top-100 re-search: ms (tensor fraction) | %.1f (%.2f) | %.1f (%.2f) | %.2f (%.2f) |' % sum(((d['alpha_qe']['ms'], d['alpha_qe']['roofline']['frac']) for d in (d1, d2, d8)), ()))
L.append('| [4] DBA, one FULL pass over the 1M rows: s (tensor fraction) | %.2f (%.2f) | %.2f (%.2f) | %.3f (%.2f) |' % sum(((d['dba']['ms']/1e3, d['dba']['roofline']['frac']) for d in (d1, d2, d8)), ()))
L.append('| [0] ResNet50-GeM extract at 1024 px, images/s (stock backbone, host images) | %.0f (CPU port: %.1f) | %.0f | %.0f |' % (d1['extract_rank']['images_per_s'], d1['extract_rank']['cpu_baseline']['value'], d2['extract_rank']['images_per_s'], d8['extract_rank']['images_per_s']))
L.append('| [0] 70 × 4,993 full ranking / top-100 | %.3f / %.3f ms | | |' % (d1['extract_rank']['full_ranking_ms'], d1['extract_rank']['top100_ms']))
L.append('| parity object (sharded == single / brute force) | n/a | all true: %s | all true: %s |' % (d2['parity']['all_ok'], d8['parity']['all_ok']))
L.append('| CPU baselines (oracle port, %d host threads) | tail %.0f desc/s; search %.1f q/s | | |' % (d1['cpu_baseline']['cores'], d1['cpu_baseline']['value'], d1['search']['cpu_baseline']['value']))
print('\n'.join(L))
