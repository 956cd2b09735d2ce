"""Optimal-transport fingering reward on the device (csrc/rp_task_kernels.cu, launched after the step kernel when
the fingering reward is disabled: piano_with_shadow_hands.py:130-141, 333-369) against the restated reference
logic (oracle/task_ref.py, scipy's linear_sum_assignment).  The solver and the reward are pinned to scipy on the host
(tests/test_hd_math.py); this covers the kernel wrapper (indexing of sites, key poses, goals)."""
import numpy as np
import pytest
import torch

import helpers
from oracle import task_ref


def _dev(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("song", ["TwinkleTwinkleRousseau", "CMajorScaleTwoHands"])
def test_ot_fingering_reward_matches_reference_logic(song):
    fm, model, batch = helpers.make_native("piano_with_shadow_hands", song=song, nenv=3, disable_fingering_reward=True,
                                           ot_fingering_reward=True)
    goal, fing = helpers.song_tables(song)
    ref = task_ref.RefEnv(fm, goal, fing, n_steps_lookahead=1, disable_fingering_reward=True, ot_fingering_reward=True)
    N = 3
    obs = torch.empty((N, model.obs_dim), dtype=torch.float64, device="cuda")
    rew = torch.empty(N, dtype=torch.float64, device="cuda"); dis = torch.empty_like(rew)
    st = torch.empty(N, dtype=torch.uint8, device="cuda")
    terms = torch.empty((N, 5), dtype=torch.float64, device="cuda")
    batch.reset(None, obs.data_ptr()); ref.reset()
    acts = helpers.twinkle_actions()
    seen_partial = 0
    for t in range(goal.shape[0] + 2):      # through the end of the episode and the auto-reset
        a = acts[t % 158]
        batch.step(_dev(np.tile(a, (N, 1))).data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), st.data_ptr(), terms.data_ptr())
        rst, rrew, _, robs = ref.step(a)
        torch.cuda.synchronize()
        assert int(st[0]) == rst
        if rst == task_ref.FIRST:
            assert float(rew[0]) == 0.0
            continue
        r_ot = ref.reward_terms["ot_fingering_reward"]
        assert abs(float(terms[0, 3]) - r_ot) < 1e-9 and abs(float(rew[0]) - rrew) < 1e-9, (t, float(terms[0, 3]), r_ot)
        assert torch.equal(terms[0], terms[2]) and abs(float(terms[0].sum()) - float(rew[0])) < 1e-12
        assert np.abs(obs[0].cpu().numpy() - robs).max() < 1e-9
        seen_partial += 0.0 < r_ot < 1.0
    assert seen_partial > 10
    batch.close()
