"""Import stand-in (TEST INFRASTRUCTURE): MS-SSIM is a third-party metric outside the hot path; returns 0."""
import torch


def ms_ssim(a, b, data_range=1.0, size_average=True):
    return torch.zeros(())
