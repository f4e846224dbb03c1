import sys, os, json
sys.path.insert(0, '/root/repo')
import interpolation_engine_b200 as ie
if len(sys.argv) > 1: ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), sys.argv[1])
import numpy as np, torch
from interpolation_engine_b200 import workloads
import bench_aux
from tests import oracle_lib
eng = ie.Engine(0)
line = bench_aux.bench_c3(eng, ie, workloads, torch, torch.device('cuda', 0), oracle_lib.load())
print(sys.argv[1:] , 'c3', line['value'], line['ms_per_step'], 'general', line['config'].get('general_path_templates'))
