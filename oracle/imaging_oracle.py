"""CPU restatement of the visibility front-end of the reference's acoustic imaging. TEST INFRASTRUCTURE (tests only).

Follows audiblelight/imaging.py: extract_visibilities :455-492, form_visibility :697-719. scikit-image's
view_as_blocks / view_as_windows (not installed here) are reshapes of contiguous frames and are written as such.
Pinned by tests/golden/imaging.npz, produced by the unmodified reference functions (tests/golden/make_golden_imaging.py).
"""
import numpy as np
from scipy.signal import windows


def extract_visibilities(data_, rate_, t, fc, bw, alpha):
    n_stft_sample = int(rate_ * t)
    if n_stft_sample == 0:
        raise ValueError("Not enough samples per time frame.")
    n_sample = (data_.shape[0] // n_stft_sample) * n_stft_sample
    n_channel = data_.shape[1]
    stf_data = data_[:n_sample].reshape(-1, n_stft_sample, n_channel)           # view_as_blocks(...).squeeze(axis=1)
    window = windows.tukey(M=n_stft_sample, alpha=alpha, sym=True).reshape(1, -1, 1)
    stf_win_data = stf_data * window
    n_stf = stf_win_data.shape[0]
    stft_data = np.fft.fft(stf_win_data, axis=1)
    idx_start = int((fc - 0.5 * bw) * n_stft_sample / rate_)
    idx_end = int((fc + 0.5 * bw) * n_stft_sample / rate_)
    collapsed = np.sum(stft_data[:, idx_start:idx_end + 1, :], axis=1)
    return collapsed.reshape(n_stf, -1, 1).conj() * collapsed.reshape(n_stf, 1, -1)


def form_visibility(data, rate, fc, bw, t_sti, t_stationarity):
    s_sti = extract_visibilities(data, rate, t_sti, fc, bw, alpha=1.0)
    n_channel = data.shape[1]
    per = int(t_stationarity / t_sti)
    n_blocks = s_sti.shape[0] // per                                               # view_as_windows, step == window
    return s_sti[:n_blocks * per].reshape(n_blocks, per, n_channel, n_channel).sum(axis=1)
