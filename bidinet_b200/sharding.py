"""Host-side logic of the multi-GPU path: agents are independent (no cross-agent term
anywhere in neo/Agent.cpp), so N GPUs = N engines, each owning a contiguous block of
agents.  There is no data-path collective; the only communication is the barrier and the
max-over-ranks of the step time that bench.py reports."""
import os

import numpy as np

N_STATE, N_CLOCK, INPUT_COUNT, N_ACTION = 15, 4, 20, 10  # RunnerMain.cpp:115-116
FB_IDX = np.arange(INPUT_COUNT, INPUT_COUNT + N_ACTION, dtype=np.int32)
SEED_BASE = 1234  # global agent g is created from std::mt19937(SEED_BASE + g)


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment (1 process: 0, 1, 0)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def agent_block(rank, agents_per_gpu):
    """First global agent index and count owned by `rank` (weak scaling: fixed per GPU)."""
    return rank * agents_per_gpu, agents_per_gpu


def seed_base_for(rank, agents_per_gpu):
    """seed_base to pass to bidi_init_random so that local agent a is global agent first+a."""
    return SEED_BASE + agent_block(rank, agents_per_gpu)[0]


def base_frames(n_frames, agents, first_agent, n_in=64, t0=0):
    """Open-loop part of the headless RunnerMain input, [T][A][n_in]:
    15 sine state values, the four clock inputs of RunnerMain.cpp:235-236; the action
    slots 20..29 are filled by the feedback loop."""
    t = (t0 + np.arange(n_frames, dtype=np.float32))[:, None, None]
    a = (first_agent + np.arange(agents, dtype=np.float32))[None, :, None]
    f = np.zeros((n_frames, agents, n_in), dtype=np.float32)
    i = np.arange(N_STATE, dtype=np.float32)[None, None, :]
    f[:, :, :N_STATE] = 0.5 + 0.5 * np.sin(0.05 * t * (1.0 + i) + 0.1 * a)
    c = np.arange(N_CLOCK, dtype=np.float32)[None, None, :]
    f[:, :, N_STATE:N_STATE + N_CLOCK] = np.sin(t / 60.0 * 2.0 * c * 2.0 * 3.141596) * 0.5 + 0.5
    return f


def max_over_ranks(value, dist=None, device=None):
    """The slowest rank's value (what bench.py reports as the step time)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(agents_per_gpu, world, steps, seconds):
    """agent-steps/s of the whole job: all ranks' agents over the slowest rank's time."""
    return world * agents_per_gpu * steps / seconds
