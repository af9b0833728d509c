"""Minimal experiment logger with the reference's interface (utils/logger.py:246-364): ``logkv``, ``logkv_mean``,
``dumpkvs``, ``set_timestep``, ``log``, ``log_hyperparameters``, directory attributes and ``make_log_dirs``.
Host-side I/O only (stdout + csv); TensorBoard output of the reference is not reproduced."""
import csv
import datetime
import json
import os
from collections import defaultdict
from typing import Dict, Iterable, Optional

ROOT_DIR = "log"


def make_log_dirs(task_name: str, algo_name: str, seed, args: Dict, part: Optional[str] = None,
                  record_params: Optional[Iterable[str]] = None) -> str:
    """utils/logger.py:346-365: log/<task>/<algo>[&param=value...]/[part/]timestamp_<stamp>&<seed>."""
    if record_params is not None:
        algo_name += "".join(f"&{p}={args[p]}" for p in record_params)
    stamp = datetime.datetime.now().strftime("%y-%m%d-%H%M%S")
    exp_name = f"timestamp_{stamp}&{seed}"
    parts = [ROOT_DIR, task_name, algo_name] + ([part] if part is not None else []) + [exp_name]
    path = os.path.join(*parts)
    os.makedirs(path, exist_ok=True)
    return path


class _CsvStream:
    """One ``<name>.csv`` progress stream (reference: utils/logger.py:144-198 CSVOutputHandler): the header grows when
    new keys appear (the file is re-emitted with the wider header), missing values stay empty."""

    def __init__(self, path: str, name: str) -> None:
        self.path, self.handler_name, self.keys, self.rows = path, name, [], []

    def writekvs(self, kvs: Dict) -> None:
        extra = sorted(k for k in kvs if k not in self.keys)
        self.rows.append(dict(kvs))
        if extra or not os.path.exists(self.path):
            self.keys.extend(extra)
            with open(self.path, "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(self.keys)
                for r in self.rows:
                    w.writerow([r.get(k, "") for k in self.keys])
        else:
            with open(self.path, "a", newline="") as f:
                csv.writer(f).writerow([kvs.get(k, "") for k in self.keys])


class Logger:
    def __init__(self, dir: str, ouput_config: Optional[Dict] = None) -> None:
        self._dir = dir
        self._name2val = defaultdict(float)
        self._name2cnt = defaultdict(int)
        self._timestep = 0
        for sub in ("record", "checkpoint", "model", "result"):
            os.makedirs(os.path.join(dir, sub), exist_ok=True)
        # ``ouput_config`` = {file name: "csv" | "tensorboard"} as in the run scripts (run_mopo.py:224-229: the streams
        # "dynamics_training_progress" and "policy_training_progress" are told apart by dumpkvs(exclude=...));
        # TensorBoard streams are not reproduced.
        cfg = ouput_config if ouput_config is not None else {"progress": "csv"}
        self._streams = [_CsvStream(os.path.join(self.record_dir, f"{name}.csv"), name)
                         for name, fmt in cfg.items() if fmt == "csv"]
        self.quiet = False

    record_dir = property(lambda self: os.path.join(self._dir, "record"))
    checkpoint_dir = property(lambda self: os.path.join(self._dir, "checkpoint"))
    model_dir = property(lambda self: os.path.join(self._dir, "model"))
    result_dir = property(lambda self: os.path.join(self._dir, "result"))

    def log_hyperparameters(self, hyper_param: Dict) -> None:
        with open(os.path.join(self.record_dir, "hyper_param.json"), "w") as f:
            json.dump({k: (v if isinstance(v, (int, float, str, bool, list, type(None))) else str(v))
                       for k, v in hyper_param.items()}, f, indent=2)

    def logkv(self, key, val) -> None:
        self._name2val[key] = val

    def logkv_mean(self, key, val) -> None:
        old, cnt = self._name2val[key], self._name2cnt[key]
        self._name2val[key] = old * cnt / (cnt + 1) + val / (cnt + 1)
        self._name2cnt[key] = cnt + 1

    def set_timestep(self, timestep: int) -> None:
        self._timestep = timestep

    def dumpkvs(self, exclude=None) -> None:
        """utils/logger.py:300-309: every stream not named in ``exclude`` gets the row."""
        row = {**self._name2val, "timestep": self._timestep}
        if not self.quiet:
            print(" | ".join(f"{k}={v:.5g}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items()), flush=True)
        for st in self._streams:
            if exclude is not None and st.handler_name in exclude:
                continue
            st.writekvs(row)
        self._name2val.clear()
        self._name2cnt.clear()

    def log(self, s: str, level=20) -> None:      # 20 = INFO, the reference's default (utils/logger.py:311)
        if not self.quiet:
            print(s, flush=True)

    def close(self) -> None:
        pass
