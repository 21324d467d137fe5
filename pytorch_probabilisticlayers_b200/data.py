"""Synthetic regression data with the reference's semantics
(reference: data/ProbabilisticLayers_SyntheticData.py:25-61).  Host-side numpy, not accelerated."""
import numpy as np
import torch


def generate_linreg_data(num_samples=1000, x_noise_std=0.01, y_noise_std=0.1, plot=False):
    x = np.linspace(-5, 5, num_samples)
    x += np.random.normal(0., y_noise_std, size=x.shape)
    y = x + np.random.normal(0., y_noise_std, size=x.shape)
    return torch.from_numpy(x.reshape(-1, 1)).float(), torch.from_numpy(y.reshape(-1, 1)).float()


def generate_nonstationary_data(num_samples=1000, x_noise_std=0.01, y_noise_std=0.1, plot=False):
    x = np.linspace(-0.35, 0.55, num_samples)
    x_noise = np.random.normal(0., x_noise_std, size=x.shape)
    y_noise = np.random.normal(loc=0, scale=np.linspace(0, y_noise_std, num_samples))
    y = x + 0.3 * np.sin(2 * np.pi * (x + x_noise)) + 0.3 * np.sin(4 * np.pi * (x + x_noise)) + y_noise
    return torch.from_numpy(x.reshape(-1, 1)).float(), torch.from_numpy(y.reshape(-1, 1)).float()
