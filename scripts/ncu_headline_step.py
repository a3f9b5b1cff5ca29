import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama3_np_b200
from llama3_np_b200 import Llama
from llama3_np_b200.config import named_config
args, hidden = named_config("stories15M", max_batch_size=256, max_seq_len=256, dtype="float32")
m = Llama(None, args, hidden_dim=hidden, random_seed=0)
ids = np.random.default_rng(1).integers(3, 32000, (256, 8))
out = m.generate_all(ids, 140)
print(out.shape)
