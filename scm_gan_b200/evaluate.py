"""Rollout-MSE evaluation (the second half of BASELINE.json's metric): host-side mirror of the reference's
`measure_prediction_mse` (main.py:784-836) on the drop-in modules.

Same arithmetic, but nothing is read back per step: the reference calls `float()` on four scalars for every timestep
(one device->host synchronisation each, main.py:817-829); here the per-step statistics stay on the device and come
back in ONE copy at the end.  The reference's early exit when every trajectory is done is applied afterwards.
"""
import torch


@torch.no_grad()
def measure_prediction_mse(nets, states, rewards, dones, actions):
    """states [B,T,C,H,W] f32, rewards [B,T,R] f32, dones [B,T] f32, actions [B,T] int64 (device tensors).
    Returns (mse_losses, mse_stddevs, reward_losses, reward_stddevs): python lists, one entry per t = 2..T-1."""
    enc, dec, rew, tr = nets["encoder"], nets["decoder"], nets["reward_predictor"], nets["transition"]
    was_training = [m.training for m in (enc, dec, rew, tr)]
    for m in (enc, dec, rew, tr):
        m.eval()
    try:
        B, T = states.shape[0], states.shape[1]
        A = tr.conv1.module.weight_bar.shape[1] - tr.latent_size
        eye = torch.eye(A, dtype=torch.float32, device=states.device)
        z = enc(states[:, :3])
        z = tr(z, eye[actions[:, 1]])
        mask = torch.ones(B, dtype=torch.float32, device=states.device)
        rows = []
        for t in range(2, T):
            mask = mask * (1 - dones[:, t])
            live = torch.sum(mask)
            predicted = torch.sigmoid(dec(z))
            diffs = mask * ((states[:, t] - predicted) ** 2).mean(dim=-1).mean(dim=-1).mean(dim=-1)
            r_diffs = mask * (rewards[:, t].sum(-1) - rew(z).sum(-1)) ** 2
            scale = B / live
            rows.append(torch.stack([torch.mean(diffs) * scale, torch.std(diffs) * scale,
                                     torch.mean(r_diffs) * scale, torch.std(r_diffs) * scale, live]))
            z = tr(z, eye[actions[:, t]])
        table = torch.stack(rows).cpu()  # the only device->host transfer
    finally:
        for m, w in zip((enc, dec, rew, tr), was_training):
            m.train(w)
    out = ([], [], [], [])
    for row in table.tolist():
        if row[4] == 0:  # main.py:809-811: stop at the maximum trajectory length
            break
        for lst, v in zip(out, row[:4]):
            lst.append(v)
    return out
