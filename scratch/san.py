import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, helpers
fm, model, batch = helpers.make_native("piano_with_shadow_hands", nenv=2)
acts = helpers.twinkle_actions()
ct = torch.tensor(np.tile(acts[0,:44], (2, 1)), dtype=torch.float64, device="cuda")
for t in range(3):
    batch.physics_step(ct.data_ptr(), 1)
torch.cuda.synchronize()
print("ok", batch.get_state()[0][0,88:96])
