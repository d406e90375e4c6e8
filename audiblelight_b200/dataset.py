"""Batch / dataset driver (scope row f2): what `Scene.generate` (core.py:1789-1874) and the loop of
`scripts/seld/generate_dataset.py:290-376` do for MANY scenes, with the audio of a whole batch rendered, mixed and
packed to 16-bit PCM by one GPU call.

    from audiblelight_b200 import dataset
    dataset.generate_scenes(scenes, output_dir, audio_fnames=[...], metadata_fnames=[...])

per scene this writes, with the reference's file naming,
    <audio_fname>_<mic>.wav      16-bit PCM, (T, C) interleaved   (sf.write(path, mix.T, sr), core.py:1840-1847)
    <metadata_fname>.json        scene.to_dict()                   (core.py:1856-1862)
    <metadata_fname>_<mic>.csv   DCASE 2024 rows                   (synthesize.py:742-878, core.py:1865-1874)

`generate_dcase2024_metadata` below restates synthesize.py:742-878; it is integer host code and is pinned BIT-EXACTLY
by the reference's own expected tables (tests/test_dcase_metadata.py:247-352) and by golden rows produced by the
unmodified reference function (tests/golden/dcase.npz). The WAV container follows libsndfile's layout for a plain
PCM_16 WAV (44-byte header); soundfile is not installable offline, so the container bytes are checked by reading the
files back, not against soundfile itself.
"""
from __future__ import annotations

import json
import os
import struct
from collections import Counter
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

DCASE_2024_COLUMNS = ["frame_number", "active_class_index", "source_number_index", "azimuth", "elevation", "distance"]


# ---- DCASE 2024 metadata ----------------------------------------------------------------------------------------------
def _frame_index(frames: np.ndarray, t: float) -> int:
    # np.where(frames == round(t, 1))[0][0] of synthesize.py:793-796 (an IndexError if t does not sit on the grid)
    return int(np.where(frames == round(t, 1))[0][0])


def dcase2024_rows(scene, temporal_resolution=0.1) -> Dict[str, np.ndarray]:
    """{mic_alias: int64 array (n_rows, 6)} in the column order of DCASE_2024_COLUMNS, sorted by (frame, class,
    source) — the content of the DataFrames `generate_dcase2024_metadata` returns (synthesize.py:742-878).

    Frames are 100 ms (`temporal_resolution`); an event is active on every frame from round(start, 1) to
    round(end, 1) inclusive; static events repeat the rounded polar position of their single emitter, moving events
    interpolate the emitters' positions linearly over the event's frames (np.interp on a linspace of the same span);
    azimuth / elevation are rounded degrees, distance is rounded centimetres (Python round, half to even); the source
    index counts per class in order of scene_start, events sharing an audio file share the index."""
    frames = np.round(np.arange(0, scene.duration + temporal_resolution, temporal_resolution), 1)
    mics = list(scene.state.microphones.keys())
    rows: Dict[str, List[List[int]]] = {m: [] for m in mics}
    per_class = Counter()
    source_of_file = {}
    for event in sorted(scene.get_events(), key=lambda e: e.scene_start):
        first = _frame_index(frames, max(event.scene_start, 0.0))
        last = _frame_index(frames, min(event.scene_end, scene.duration))
        active = np.arange(first, last + 1)
        if not isinstance(event.class_id, int):
            raise ValueError("Can't convert Event to DCASE format without valid DCASE class indices")
        if event.filename not in source_of_file:
            source_of_file[event.filename] = per_class.get(event.class_id, 0)
            per_class[event.class_id] += 1
        source = source_of_file[event.filename]
        for mic in mics:
            if not event.is_moving:
                az, el, dist = event.emitters[0].coordinates_relative_polar[mic][0]
                az, el, dist = round(az), round(el), round(dist * 100)
                rows[mic] += [[int(f), event.class_id, source, az, el, dist] for f in active]
            else:
                coords = np.vstack([e.coordinates_relative_polar[mic] for e in event.emitters])
                t_frames = frames[active]
                t_coords = np.linspace(min(t_frames), max(t_frames), num=len(coords))
                track = np.stack([np.interp(t_frames, t_coords, coords[:, d]) for d in range(coords.shape[1])], axis=1)
                for f, (az, el, dist) in zip(active, track):
                    rows[mic].append([int(f), event.class_id, source, round(az), round(el), round(dist * 100)])
    out = {}
    for mic, data in rows.items():
        a = np.asarray(data, dtype=np.int64).reshape(-1, 6)
        order = np.lexsort((a[:, 2], a[:, 1], a[:, 0]))  # stable, like DataFrame.sort_values on the three keys
        out[mic] = a[order]
    return out


def generate_dcase2024_metadata(scene, temporal_resolution=0.1):
    """Same return value as the reference function (synthesize.py:742-878): {mic: DataFrame indexed by
    frame_number}. Needs pandas; `dcase2024_rows` / `dcase_csv` do not."""
    import pandas as pd
    return {mic: pd.DataFrame(a, columns=DCASE_2024_COLUMNS).set_index("frame_number")
            for mic, a in dcase2024_rows(scene, temporal_resolution).items()}


def dcase_csv(rows: np.ndarray) -> str:
    """Text of `df.to_csv(path, sep=",", encoding="utf-8", header=None)` (core.py:1874) for one microphone."""
    return "".join(",".join(str(int(v)) for v in r) + os.linesep for r in np.asarray(rows).reshape(-1, 6))


# ---- WAV (PCM_16) ------------------------------------------------------------------------------------------------------
def write_wav_pcm16(path: Union[str, Path], pcm: np.ndarray, sample_rate: int) -> None:
    """(T, C) int16 -> RIFF/WAVE file with the canonical 44-byte PCM header (what libsndfile emits for
    sf.write(path, float_data, sr) on a .wav path: format tag 1, 16 bits, little endian)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    n_frames, n_ch = pcm.shape
    data_bytes = n_frames * n_ch * 2
    if data_bytes > 0xFFFFFFFF - 36:
        raise ValueError("audio too long for a RIFF container")
    sr = int(sample_rate)
    header = b"RIFF" + struct.pack("<I", 36 + data_bytes) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, n_ch, sr, sr * n_ch * 2, n_ch * 2, 16) + b"data" + struct.pack("<I", data_bytes)
    with open(path, "wb") as f:
        f.write(header)
        f.write(memoryview(pcm).cast("B"))


def read_wav_pcm16(path: Union[str, Path]):
    """(sample_rate, (T, C) int16) of a file written by `write_wav_pcm16`."""
    with open(path, "rb") as f:
        b = f.read()
    if b[:4] != b"RIFF" or b[8:16] != b"WAVEfmt ":
        raise ValueError("not a PCM WAV file")
    _, tag, n_ch, sr, _, _, bits = struct.unpack("<IHHIIHH", b[16:36])
    if tag != 1 or bits != 16 or b[36:40] != b"data":
        raise ValueError("not a 16-bit PCM WAV file")
    n = struct.unpack("<I", b[40:44])[0]
    return sr, np.frombuffer(b, dtype="<i2", count=n // 2, offset=44).reshape(-1, n_ch)


# ---- the driver --------------------------------------------------------------------------------------------------------
def _with_mic(path: Path, mic: str, suffix: str) -> Path:
    # path.with_suffix(".wav").with_stem(f"{path.name}_{mic}") of core.py:1841-1845
    return path.parent / f"{path.name}_{mic}{suffix}"


def generate_scenes(scenes: Sequence, output_dir: Optional[Union[str, Path]] = None, audio: bool = True,
                    metadata_json: bool = True, metadata_dcase: bool = True,
                    audio_fnames: Optional[Sequence[Union[str, Path]]] = None,
                    metadata_fnames: Optional[Sequence[Union[str, Path]]] = None,
                    batch_scenes: int = 16, device: int = -1, keep_audio: bool = False,
                    pipeline: bool = True) -> List[Dict[str, List[Path]]]:
    """`scene.generate(output_dir, audio, metadata_json, metadata_dcase, audio_fname, metadata_fname)` for every scene
    of the list, `batch_scenes` scenes per GPU call. File names default to audio_out_<i> / metadata_out_<i>.
    `keep_audio=True` also leaves `scene.audio` / `event.spatial_audio` populated as `Scene.generate` would
    (more device->host traffic). With `pipeline=True` batch i+1 is prepared and rendered (second context, worker
    thread; the library call releases the GIL) while the files of batch i are written. Returns the paths written,
    per scene."""
    from concurrent.futures import ThreadPoolExecutor

    from . import synthesize as _syn
    n = len(scenes)
    out_dir = Path(output_dir) if output_dir is not None else Path.cwd()
    if not out_dir.is_dir():
        raise FileNotFoundError(f"Output directory {out_dir} does not exist")
    audio_fnames = list(audio_fnames) if audio_fnames is not None else [f"audio_out_{i:04d}" for i in range(n)]
    metadata_fnames = list(metadata_fnames) if metadata_fnames is not None else [f"metadata_out_{i:04d}" for i in range(n)]
    if len(audio_fnames) != n or len(metadata_fnames) != n:
        raise ValueError("need one audio / metadata file name per scene")
    written: List[Dict[str, List[Path]]] = [dict(audio=[], json=[], csv=[]) for _ in range(n)]
    step = max(1, int(batch_scenes))
    starts = list(range(0, n, step))

    def render(b0, renderer):
        batch = list(scenes[b0:b0 + step])
        if not audio:
            return batch, None
        return batch, _syn.render_scenes(batch, ignore_cache=True, device=device, store_padded=keep_audio, pcm16=True,
                                         keep_event_audio=keep_audio, keep_mix=keep_audio, renderer=renderer,
                                         pinned=True)  # the PCM arrays are written to disk before the pool is reused

    def write(b0, batch, pcm):
        for k, scene in enumerate(batch):
            if pcm is not None:
                base = (out_dir / audio_fnames[b0 + k]).with_suffix("")
                for mic, data in pcm[k].items():
                    p = _with_mic(base, mic, ".wav")
                    write_wav_pcm16(p, data, int(scene.sample_rate))
                    written[b0 + k]["audio"].append(p)
            base = (out_dir / metadata_fnames[b0 + k]).with_suffix("")
            if metadata_json:
                p = base.with_suffix(".json")
                with open(p, "w") as f:
                    json.dump(scene.to_dict(), f, indent=4, ensure_ascii=False)
                written[b0 + k]["json"].append(p)
            if metadata_dcase:
                for mic, rows in dcase2024_rows(scene).items():
                    p = _with_mic(base, mic, ".csv")
                    with open(p, "w", encoding="utf-8", newline="") as f:
                        f.write(dcase_csv(rows))
                    written[b0 + k]["csv"].append(p)

    if not pipeline or not audio or len(starts) < 2:
        for b0 in starts:
            write(b0, *render(b0, None))
        return written
    # one persistent context per batch in flight (their workspaces and pinned pools stay warm between calls)
    contexts = [_syn.get_renderer(device, 0), _syn.get_renderer(device, 1)]
    with ThreadPoolExecutor(max_workers=1) as pool:
        pending = pool.submit(render, starts[0], contexts[0])
        for i, b0 in enumerate(starts):
            batch, pcm = pending.result()
            if i + 1 < len(starts):
                pending = pool.submit(render, starts[i + 1], contexts[(i + 1) & 1])
            write(b0, batch, pcm)
    return written
