"""Checkpoint interop with the reference (reference: Learner.save_checkpoint / resume_checkpoint /
load_checkpoint_best / load_checkpoint_epoch / ensembling / load_checkpoint_ensemble / remove_checkpoint_epochs,
code/learner.py:302-486).  Same file names (`latest_model.tar`, `best_model.tar`, `model{N}.tar`, `ensemble_model.tar`) and the
same dictionary layout ({"epoch", "max_score", "model"[, "scaler"]}), so checkpoints flow both ways.  Checkpoints written
through nn.DataParallel carry a `module.` prefix on every key (learner.py:30-31); it is stripped on load."""
import os

import torch


def _path(checkpoints_dir, name):
    return os.path.join(checkpoints_dir, name)


def _strip_dataparallel(sd):
    if sd and all(k.startswith("module.") for k in sd):
        return {k[len("module."):]: v for k, v in sd.items()}
    return sd


def _read(path, device):
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} does not exist, can not load the checkpoint")
    ck = torch.load(path, map_location=device, weights_only=False)
    ck["model"] = _strip_dataparallel(ck["model"])
    return ck


def _load_into(model, sd, as_all_state, ex_key):
    """as_all_state: strict load.  Otherwise copy every entry whose (ex_key + key) exists in the model (partial / prefixed load)."""
    if as_all_state:
        model.load_state_dict(sd)
        return set(sd)
    own = model.state_dict()
    matched = {k for k in sd if ex_key + k in own}
    if len(matched) <= 1:
        raise ValueError("loaded model parameters and original parameters unmatched")
    for k in matched:
        own[ex_key + k] = sd[k]
    model.load_state_dict(own)
    return matched


class CheckpointMixin:
    """Mixed into Learner: needs self.model, self.max_score, self.start_epoch, self.use_amp, self.device (+ self.scaler under amp)."""

    def early_stopping(self, current_score, patience=5):
        """learner.py:283-300 -> (stop_flag, is_best_epoch)"""
        if self.is_best_epoch(current_score):
            self.early_stop_counter = 0
            return False, True
        self.early_stop_counter += 1
        return self.early_stop_counter >= patience, False

    def is_best_epoch(self, current_score):
        """learner.py:333-342"""
        if current_score >= self.max_score:
            self.max_score = current_score
            return True
        return False

    def save_checkpoint(self, epoch, checkpoints_dir, is_best_epoch=False, save_extra_hist=False, dataparallel_keys=False):
        """learner.py:344-374.  dataparallel_keys=True writes every key with the `module.` prefix an nn.DataParallel-wrapped reference
        model expects (learner.py:28-31: after mul_gpu() the reference saves and loads `module.`-prefixed state_dicts); the default
        writes the plain keys a single-GPU reference run reads.  Our own loaders accept both."""
        sd = self.model.state_dict()
        if dataparallel_keys:
            sd = {"module." + k: v for k, v in sd.items()}
        state = {"epoch": epoch, "max_score": self.max_score, "model": sd}
        if self.use_amp:
            state["scaler"] = self.scaler.state_dict()
        torch.save(state, _path(checkpoints_dir, "latest_model.tar"))
        if save_extra_hist:
            torch.save(state, _path(checkpoints_dir, f"model{epoch}.tar"))
        if is_best_epoch:
            torch.save(state, _path(checkpoints_dir, "best_model.tar"))

    def resume_checkpoint(self, checkpoints_dir, from_latest=True, as_all_state=True, ex_key=""):
        """learner.py:377-408"""
        ck = _read(_path(checkpoints_dir, "latest_model.tar" if from_latest else "best_model.tar"), self.device)
        self.start_epoch = ck["epoch"] + 1
        self.max_score = ck["max_score"]
        if self.use_amp and "scaler" in ck:
            self.scaler.load_state_dict(ck["scaler"])
        _load_into(self.model, ck["model"], as_all_state, ex_key)

    def load_checkpoint_best(self, checkpoints_dir, as_all_state=True, param_frozen=False, ex_key=""):
        """learner.py:414-448 -> epoch of the best model"""
        ck = _read(_path(checkpoints_dir, "best_model.tar"), self.device)
        matched = _load_into(self.model, ck["model"], as_all_state, ex_key)
        if param_frozen:
            for name, p in self.model.named_parameters():
                if name in matched or (ex_key and name.startswith(ex_key) and name[len(ex_key):] in matched):
                    p.requires_grad = False
        return ck["epoch"]

    def load_checkpoint_epoch(self, checkpoints_dir, epoch):
        """learner.py:450-465"""
        ck = _read(_path(checkpoints_dir, f"model{epoch}.tar"), self.device)
        if ck["epoch"] != epoch:
            raise ValueError(f"checkpoint holds epoch {ck['epoch']}, expected {epoch}")
        self.model.load_state_dict(ck["model"])

    def ensembling(self, checkpoints_dir, epochs):
        """learner.py:302-331: average the state_dicts of `epochs` (every entry, like the reference), load and save the result."""
        acc = None
        for e in epochs:
            sd = _read(_path(checkpoints_dir, f"model{e}.tar"), self.device)["model"]
            if acc is None:
                acc = {k: v * 1 / len(epochs) for k, v in sd.items()}
            else:
                for k, v in sd.items():
                    acc[k] += v * 1 / len(epochs)
        own = self.model.state_dict()
        acc = {k: v.to(own[k].dtype) for k, v in acc.items()}
        self.model.load_state_dict(acc)
        torch.save({"epoch": epochs, "model": self.model.state_dict()}, _path(checkpoints_dir, "ensemble_model.tar"))

    def load_checkpoint_ensemble(self, checkpoints_dir):
        """learner.py:467-479"""
        ck = _read(_path(checkpoints_dir, "ensemble_model.tar"), self.device)
        self.model.load_state_dict(ck["model"])
        return ck["epoch"]

    def remove_checkpoint_epochs(self, checkpoints_dir, epochs):
        """learner.py:481-486"""
        for e in epochs:
            os.remove(_path(checkpoints_dir, f"model{e}.tar"))
