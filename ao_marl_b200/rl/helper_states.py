"""Integer index tables that say which state entries each agent sees and where its actions go.

Restates (bit-exact; checked against tests/golden/ref_rl.npz) the reference's
  create_agents_dictionary_original   src/.../rpc_training/train_rpc.py:265-296
  select_correct_modes_for_array      src/.../helper_rpc/helper_states.py:7-27
  modes_chosen_original               helper_states.py:30-69
  modes_chosen_window_n_zernike       helper_states.py:202-283
  get_modes_chosen                    helper_states.py:286-317
  get_state_shape_worker              train_rpc.py:310-334
"""
import numpy as np


def agents_dictionary(total_existing_modes, start, end, world_size, include_tip_tilt):
    """worker id (1-based) -> [first mode, last mode + 1]; the tip-tilt agent, if any, comes last."""
    controlled = end - start
    if start < 0 or end < 0 or controlled <= 0:
        raise ValueError("n_zernike_start_end must be set for the multi-agent layout")
    n_mode_agents = world_size - 2 if include_tip_tilt else world_size - 1
    if n_mode_agents <= 0 or controlled % n_mode_agents:
        raise ValueError("controlled modes must divide evenly among the agents")
    per = controlled // n_mode_agents
    d = {}
    wid = 1
    for m in range(start, end, per):
        d[wid] = [m, m + per]
        wid += 1
    if include_tip_tilt:
        d[wid] = [total_existing_modes - 2, total_existing_modes]
        controlled += 2
    return d, controlled, per


def action_slots(agent_range, total_existing_modes, total_controlled_modes, starting_mode, include_tip_tilt):
    """[bottom, top) of one agent inside the global action vector (TT agent = last two entries)."""
    a0, a1 = agent_range
    if include_tip_tilt and a0 == total_existing_modes - 2 and a1 == total_existing_modes:
        return total_controlled_modes - 2 - starting_mode, total_controlled_modes - starting_mode
    return a0 - starting_mode, a1 - starting_mode


def modes_chosen_plain(agents, indices_of_state, total_existing_modes, total_controlled_modes, starting_mode,
                       include_tip_tilt):
    out = {}
    for wid, rng in agents.items():
        lo, hi = action_slots(rng, total_existing_modes, total_controlled_modes, starting_mode, include_tip_tilt)
        parts = []
        for key, (k0, k1) in indices_of_state.items():
            if key in "wfs":
                parts.append(np.arange(k0, k1))
            else:
                parts.append(np.arange(k0 + lo, k0 + hi))
        out[wid] = np.concatenate(parts)
    return out


def modes_chosen_windowed(agents, indices_of_state, window, include_tip_tilt, include_tip_tilt_windowed,
                          n_filtered):
    """Each mode agent sees its own modes plus `window` neighbours on both sides, the window sliding
    inwards at the two ends of the (unfiltered, non-TT) mode range; the TT agent sees the last two
    entries of every block (plus the first 2*window when include_tip_tilt_windowed)."""
    out = {}
    n_agents = len(agents)
    for wid, (lo, hi) in agents.items():
        parts = []
        if include_tip_tilt and wid == n_agents:
            for key, (k0, k1) in indices_of_state.items():
                if key in "wfs":
                    parts.append(np.arange(k0, k1))
                else:
                    sel = np.arange(k1 - 2, k1)
                    if include_tip_tilt_windowed:
                        sel = np.concatenate([sel, np.arange(0, int(2 * window))])
                    parts.append(sel)
            out[wid] = np.concatenate(parts)
            break
        for key, (k0, k1) in indices_of_state.items():
            if key in "wfs":
                parts.append(np.arange(k0, k1))
                continue
            span = (k1 - (n_filtered - 2)) - k0
            if lo - window < 0:
                ini, end = 0, hi + window - (lo - window)
            elif hi + window > span:
                ini, end = lo - int(window) - (hi + window - span), span
            else:
                ini, end = lo - window, hi + window
            parts.append(np.arange(k0 + ini, k0 + end))
        out[wid] = np.concatenate(parts)
    return out


def get_modes_chosen(agents, indices_of_state, env_rl, n_filtered, total_existing_modes, total_controlled_modes,
                     starting_mode):
    if env_rl["window_n_zernike"] > -1:
        return modes_chosen_windowed(agents, indices_of_state, env_rl["window_n_zernike"],
                                     env_rl["include_tip_tilt"], env_rl["include_tip_tilt_windowed"], n_filtered)
    if env_rl.get("tt_treated_as_mode"):
        raise NotImplementedError("tt_treated_as_mode layouts are outside the hot-path scope")
    return modes_chosen_plain(agents, indices_of_state, total_existing_modes, total_controlled_modes,
                              starting_mode, env_rl["include_tip_tilt"])


def state_shape_worker(agent_range, env_rl, worker_id, n_agents):
    mult = (int(env_rl["state_dm_residual"]) + int(env_rl["state_dm_after_linear"]) +
            int(env_rl["state_dm_before_linear"]) + env_rl["number_of_previous_dm"] +
            env_rl["number_of_previous_dm_residuals"])
    base = (agent_range[1] - agent_range[0]) * mult
    w = env_rl["window_n_zernike"]
    if w > -1:
        if env_rl["include_tip_tilt"] and worker_id == n_agents:
            return base + (int(2 * w) * mult if env_rl["include_tip_tilt_windowed"] else 0)
        return base + int(w * 2) * mult
    return base
