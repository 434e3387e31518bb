"""Result persistence in the reference's formats (io_fwm.py): compressed .npz with keys
`z`, `A`, `metadata_json` (:73-134, loader :137-170), metadata .json (:177-212), per-sample
powers/phases .csv (:219-294) and the 3-file bundle (:297-328).  Files written here load with
the reference's `load_result_npz` and vice versa.  `save_sweep_npz` adds a batched variant for
sweep results (not in the reference)."""
from __future__ import annotations

import csv
import datetime as _dt
import json
from dataclasses import asdict, is_dataclass
from pathlib import Path
from typing import Any

import numpy as np


def _ensure_path(path) -> Path:
    return Path(path).expanduser()


def _json_default(obj: Any) -> Any:
    if is_dataclass(obj):
        return asdict(obj)
    if isinstance(obj, Path):
        return str(obj)
    if isinstance(obj, (np.integer, np.floating, np.bool_)):
        return obj.item()
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    raise TypeError(f"Object of type {type(obj).__name__} is not JSON serializable")


def _make_metadata(metadata, *, add_timestamp: bool = True) -> dict:
    md = dict(metadata) if metadata else {}
    if add_timestamp and "timestamp_utc" not in md:
        now = _dt.datetime.now(_dt.timezone.utc).replace(microsecond=0, tzinfo=None)
        md["timestamp_utc"] = now.isoformat() + "Z"
    return md


def _target(path, suffix: str, overwrite: bool) -> Path:
    p = _ensure_path(path)
    if p.suffix.lower() != suffix:
        p = p.with_suffix(suffix)
    if p.exists() and not overwrite:
        raise FileExistsError(f"File already exists: {p}")
    return p


def _check_zA(z, A, four_columns: bool = False):
    z = np.asarray(z, dtype=float)
    A = np.asarray(A)
    if z.ndim != 1:
        raise ValueError("z must be a 1D array")
    if four_columns:
        if A.ndim != 2 or A.shape[1] != 4:
            raise ValueError("A must have shape (N, 4) for this summary function")
    elif A.ndim != 2:
        raise ValueError("A must be a 2D array")
    if A.shape[0] != z.shape[0]:
        raise ValueError("A.shape[0] must match z.shape[0]")
    return z, A


def save_result_npz(path, z, A, *, metadata=None, overwrite: bool = False) -> Path:
    p = _target(path, ".npz", overwrite)
    z, A = _check_zA(z, A)
    md_json = json.dumps(_make_metadata(metadata), ensure_ascii=False, default=_json_default)
    p.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(p, z=z, A=A, metadata_json=np.array(md_json))
    return p


def load_result_npz(path) -> tuple[np.ndarray, np.ndarray, dict]:
    p = _ensure_path(path)
    if not p.exists():
        raise FileNotFoundError(f"No such file: {p}")
    with np.load(p, allow_pickle=False) as data:
        if "z" not in data or "A" not in data:
            raise ValueError("NPZ file does not contain required keys: 'z' and 'A'")
        z = np.array(data["z"], dtype=float)
        A = np.array(data["A"])
        metadata: dict = {}
        if "metadata_json" in data:
            try:
                text = str(data["metadata_json"])
                metadata = json.loads(text) if text else {}
            except Exception:
                metadata = {}
    return z, A, metadata


def save_metadata_json(path, metadata, *, overwrite: bool = False) -> Path:
    p = _target(path, ".json", overwrite)
    p.parent.mkdir(parents=True, exist_ok=True)
    with p.open("w", encoding="utf-8") as fh:
        json.dump(_make_metadata(metadata), fh, ensure_ascii=False, indent=2, default=_json_default)
    return p


def load_metadata_json(path) -> dict:
    p = _ensure_path(path)
    if not p.exists():
        raise FileNotFoundError(f"No such file: {p}")
    with p.open("r", encoding="utf-8") as fh:
        return json.load(fh)


def save_summary_csv(path, z, A, *, wave_labels=("pump 1", "pump 2", "signal", "idler"),
                     overwrite: bool = False) -> Path:
    """Columns: z, P_<label> x4, phi_<label> x4 (powers |A|^2 and phases angle(A))."""
    p = _target(path, ".csv", overwrite)
    z, A = _check_zA(z, A, four_columns=True)
    if len(wave_labels) != 4:
        raise ValueError("wave_labels must have length 4")
    table = np.column_stack((z, np.abs(A) ** 2, np.angle(A)))
    p.parent.mkdir(parents=True, exist_ok=True)
    with p.open("w", encoding="utf-8", newline="") as fh:
        writer = csv.writer(fh)
        writer.writerow(["z"] + [f"P_{w}" for w in wave_labels] + [f"phi_{w}" for w in wave_labels])
        for row in table:
            writer.writerow([float(v) for v in row])
    return p


def save_run_bundle(output_dir, run_name: str, z, A, *, metadata=None, overwrite: bool = False) -> dict:
    """<run_name>.npz + .csv + .json in output_dir; returns {'npz','csv','json': Path}."""
    out_dir = _ensure_path(output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    md = _make_metadata(metadata)
    return {
        "npz": save_result_npz(out_dir / f"{run_name}.npz", z, A, metadata=md, overwrite=overwrite),
        "csv": save_summary_csv(out_dir / f"{run_name}.csv", z, A, overwrite=overwrite),
        "json": save_metadata_json(out_dir / f"{run_name}.json", md, overwrite=overwrite),
    }


def save_sweep_npz(path, *, axes: dict, results: dict, metadata=None, overwrite: bool = False) -> Path:
    """Batched sweep output: every axis / result array under its own key + `metadata_json`
    (same 0-d unicode convention as save_result_npz)."""
    p = _target(path, ".npz", overwrite)
    md_json = json.dumps(_make_metadata(metadata), ensure_ascii=False, default=_json_default)
    arrays = {f"axis_{k}": np.asarray(v) for k, v in axes.items()}
    arrays.update({k: np.asarray(v) for k, v in results.items() if isinstance(v, np.ndarray)})
    p.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(p, metadata_json=np.array(md_json), **arrays)
    return p


def save_sweep_csv(path, *, columns: dict, overwrite: bool = False) -> Path:
    """One row per scan point: `columns` maps header -> 1-D array (all of one length), e.g.
    {"lambda3_nm": x, "gain_dB": g, "dbeta_1_m": d}.  NaN entries are written as 'nan'."""
    p = _target(path, ".csv", overwrite)
    cols = {k: np.asarray(v).reshape(-1) for k, v in columns.items()}
    sizes = {v.size for v in cols.values()}
    if len(sizes) != 1:
        raise ValueError("all columns must have the same length")
    p.parent.mkdir(parents=True, exist_ok=True)
    with p.open("w", encoding="utf-8", newline="") as fh:
        writer = csv.writer(fh)
        writer.writerow(list(cols))
        for row in zip(*cols.values()):
            writer.writerow([v.item() if hasattr(v, "item") else v for v in row])
    return p


def save_sweep_bundle(output_dir, run_name: str, *, axes: dict, results: dict, metadata=None,
                      overwrite: bool = False) -> dict:
    """Sweep analogue of save_run_bundle: <run_name>.npz (axes + result maps + metadata_json),
    <run_name>.json (metadata) and, for 1-D sweeps, <run_name>.csv (one row per scan point)."""
    out_dir = _ensure_path(output_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    md = _make_metadata(metadata)
    saved = {
        "npz": save_sweep_npz(out_dir / f"{run_name}.npz", axes=axes, results=results, metadata=md,
                              overwrite=overwrite),
        "json": save_metadata_json(out_dir / f"{run_name}.json", md, overwrite=overwrite),
    }
    flat = {k: np.asarray(v) for k, v in {**axes, **results}.items() if isinstance(v, np.ndarray)}
    one_d = {k: v for k, v in flat.items() if v.ndim == 1}
    if one_d and len({v.size for v in one_d.values()}) == 1 and len(one_d) == len(flat):
        saved["csv"] = save_sweep_csv(out_dir / f"{run_name}.csv", columns=one_d, overwrite=overwrite)
    return saved
