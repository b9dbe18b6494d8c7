import sys, json, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import bench
import vmc_jax_b200 as jVMC
import vmc_jax_b200.operator as op
print(json.dumps(bench.aux_config4_cnn(jVMC, op, torch), indent=1))
