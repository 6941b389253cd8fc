import sys, json
sys.path.insert(0, '/root/repo')
import torch, bench_plans
import global_body_planner_b200 as gbp
r = bench_plans.run_rough_k4096(gbp, torch, torch.device('cuda', 0))
print(json.dumps({k: r[k] for k in ('validated_actions_per_s', 'extends_per_s', 'solved', 'seconds')}))
