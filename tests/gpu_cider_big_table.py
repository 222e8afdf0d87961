"""GPU: the CIDEr-D reward kernel against a document-frequency table of coco-val.p size (1.6 M n-grams), checked against the
oracle on the same table and timed.  Writes gpurun_out/<tag>_cider_big_table.json."""
import json, os, sys, time
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import imagecaptioning.pytorch_b200 as b200
from oracle import ciderd_oracle as cdo

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
V, B, n, T, target = 9487, 10, 5, 20, 1_600_000
rng = np.random.RandomState(7)
gts = cdo.make_refs(B, V, seed=5)
df = {}
# n-grams of real reference-like rows first (so lookups hit), then random fill up to the coco-val.p entry count
seed_df, _ = cdo.build_document_frequency(cdo.make_refs(3000, V, seed=4) + gts)
df.update(seed_df)
while len(df) < target:
    k = rng.randint(1, 5)
    rows = rng.randint(1, V + 1, size=(200000, k))
    for r in rows:
        df[tuple(int(x) for x in r)] = float(rng.randint(1, 500))
        if len(df) >= target:
            break
ref_len = 40504.0
t0 = time.time()
table = b200.rewards.CiderDTable(df, ref_len)
build_s = time.time() - t0
sampled = np.zeros((B * n, T), np.int64)
greedy = np.zeros((B, T), np.int64)
for i in range(B * n):
    ln = rng.randint(5, T)
    sampled[i, :ln] = rng.randint(1, V + 1, size=ln)
    ref = gts[i // n][i % 5]
    sampled[i, :6] = ref[:6]                       # share n-grams with the references so the scores are not ~0
for i in range(B):
    greedy[i, :8] = gts[i][1][:8]
sd, gd = torch.from_numpy(sampled).cuda(), torch.from_numpy(greedy).cuda()
scores, reward = b200.rewards.cider_scores_and_reward(gd, gts, sd, table)
torch.cuda.synchronize()
o_reward, o_scores = cdo.self_critical_reward(greedy, gts, sampled, df, ref_len)
err_s = float(np.abs(scores.cpu().numpy() - o_scores).max())
err_r = float(np.abs(reward.cpu().numpy() - o_reward).max())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    b200.rewards.cider_scores_and_reward(gd, gts, sd, table)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
t0 = time.time()
for _ in range(3):
    cdo.self_critical_reward(greedy, gts, sampled, df, ref_len)
cpu_ms = (time.time() - t0) / 3 * 1e3
out = {'table_entries': len(df), 'ref_len': ref_len, 'table_build_s': build_s, 'hypotheses': B * n + B, 'max_abs_err_scores': err_s, 'max_abs_err_reward': err_r,
       'max_score': float(o_scores.max()), 'gpu_ms_per_reward_call_incl_ref_packing': ms, 'cpu_oracle_ms_per_call': cpu_ms}
print(json.dumps(out))
assert err_s < 1e-9 and err_r < 1e-4
os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(REPO, 'gpurun_out', '%s_cider_big_table.json' % tag), 'w'), indent=1)
