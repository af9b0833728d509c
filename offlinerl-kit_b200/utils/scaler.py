"""StandardScaler for the dynamics-model inputs (reference: utils/scaler.py:6-61); NumPy on the host, as the reference.
The device-side transform used inside ``EnsembleDynamics.step`` is ``orlk_dyn_input``."""
import os

import numpy as np
import torch


class StandardScaler:
    def __init__(self, mu=None, std=None):
        self.mu, self.std = mu, std

    def fit(self, data: np.ndarray) -> None:
        self.mu = np.mean(data, axis=0, keepdims=True)
        self.std = np.std(data, axis=0, keepdims=True)
        self.std[self.std < 1e-12] = 1.0

    def transform(self, data):
        return (data - self.mu) / self.std

    def inverse_transform(self, data):
        return self.std * data + self.mu

    def transform_tensor(self, data: torch.Tensor) -> torch.Tensor:
        return torch.tensor(self.transform(data.cpu().numpy()), device=data.device)

    def save_scaler(self, save_path: str) -> None:
        np.save(os.path.join(save_path, "mu.npy"), self.mu)
        np.save(os.path.join(save_path, "std.npy"), self.std)

    def load_scaler(self, load_path: str) -> None:
        self.mu = np.load(os.path.join(load_path, "mu.npy"))
        self.std = np.load(os.path.join(load_path, "std.npy"))
