import sys, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch, bench
print(json.dumps(bench.other_configs(torch.device("cuda", 0), only=sys.argv[1:] or None), indent=0))
