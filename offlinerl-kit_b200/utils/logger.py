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


def make_log_dirs(task_name: str, algo_name: str, seed: int, args: Dict, record_params: Optional[Iterable[str]] = None) -> str:
    if record_params is not None:
        algo_name += "".join(f"&{p}={args[p]}" for p in record_params)
    stamp = datetime.datetime.now().strftime("%y-%m%d-%H%M%S")
    path = os.path.join(ROOT_DIR, task_name, algo_name, f"seed_{seed}&timestamp_{stamp}")
    os.makedirs(path, exist_ok=True)
    return path


class Logger:
    def __init__(self, dir: str, ouput_config: Optional[Dict] = None) -> None:
        self._dir = dir
        self._name2val = defaultdict(float)
        self._name2cnt = defaultdict(int)
        self._timestep = 0
        for sub in ("record", "checkpoint", "model", "result"):
            os.makedirs(os.path.join(dir, sub), exist_ok=True)
        self._csv_path = os.path.join(self.record_dir, "progress.csv")
        self._csv_keys = None
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
        row = {"timestep": self._timestep, **self._name2val}
        if not self.quiet:
            print(" | ".join(f"{k}={v:.5g}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items()), flush=True)
        if self._csv_keys is None:
            self._csv_keys = list(row)
            with open(self._csv_path, "w", newline="") as f:
                csv.writer(f).writerow(self._csv_keys)
        with open(self._csv_path, "a", newline="") as f:
            csv.writer(f).writerow([row.get(k, "") for k in self._csv_keys])
        self._name2val.clear()
        self._name2cnt.clear()

    def log(self, s: str, level=None) -> None:
        if not self.quiet:
            print(s, flush=True)

    def close(self) -> None:
        pass
