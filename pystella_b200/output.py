"""
Run output next to the time loop: :class:`OutputFile` with the call contract of
``pystella/output.py:52-181`` -- ``output(group, **values)`` appends one record to every
dataset of a group, datasets being created on first use and all of them required
afterwards.

h5py is not part of this image, so the container is an ``.npz`` archive: key
``"<group>/<dataset>"`` holds the records stacked along a new first axis (what the
reference's resizable HDF5 datasets hold), ``"attrs/<key>"`` the run attributes.  Records
are kept in memory and the archive is rewritten atomically every ``flush_every`` records
and on :meth:`close`.  When h5py is importable, ``backend="h5py"`` writes the reference's
HDF5 layout instead.
"""

import atexit
import datetime
import os
import socket

import numpy as np

__all__ = ["OutputFile"]


class OutputFile:
    """
    :arg context: ignored (the reference stores OpenCL device information); the CUDA
        device name and driver are stored instead when a GPU is present.
    :arg name: file name without extension; default: current date and time.
    :arg runfile: a file whose content is stored as ``attrs["runfile"]``.

    Remaining keyword arguments are stored as attributes (``str(val)`` or
    ``val.__name__`` when not storable as such).
    """

    def __init__(self, context=None, name=None, runfile=None, backend=None,
                 flush_every=16, **kwargs):
        if backend not in (None, "npz", "h5py"):
            raise NotImplementedError(f"{backend} output backend")
        self.backend = backend or "npz"
        ext = ".h5" if self.backend == "h5py" else ".npz"
        stamp = name is None
        while True:
            if stamp:
                name = datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
            self.filename = name + ext
            try:
                fd = os.open(self.filename, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
                os.close(fd)
                break
            except FileExistsError:
                if not stamp:
                    raise
                import time
                time.sleep(1)
        self.flush_every = int(flush_every)
        self._records = {}
        self._pending = 0
        self.attrs = {"hostname": socket.getfqdn()}
        try:
            import torch
            if torch.cuda.is_available():
                self.attrs["device"] = torch.cuda.get_device_name()
                self.attrs["driver_version"] = str(torch.version.cuda)
        except ImportError:
            pass
        for key, val in kwargs.items():
            if isinstance(val, (tuple, list)):
                # a mixed list / tuple would become an object array, which read()
                # (allow_pickle=False) cannot load: store it as text instead
                try:
                    arr = np.asarray(val) if len(val) else np.asarray(val, dtype=np.float64)
                    plain = arr.dtype != object and arr.dtype.kind in "biufc" or (
                        arr.dtype.kind in "SU" and all(isinstance(v, str) for v in val))
                except ValueError:          # ragged
                    plain = False
                self.attrs[key] = val if plain else str(val)
            elif isinstance(val, (int, float, str, bool, np.ndarray, np.generic)):
                self.attrs[key] = val
            elif isinstance(val, type):
                self.attrs[key] = val.__name__
            else:
                self.attrs[key] = str(val)
        if runfile is not None:
            with open(runfile) as fp:
                self.attrs["runfile"] = fp.read()
        import pystella_b200
        self.attrs["versions/pystella_b200"] = pystella_b200.__version__
        self.attrs["versions/numpy"] = np.__version__
        if self.backend == "h5py":
            import h5py
            os.remove(self.filename)
            with h5py.File(self.filename, "x") as f:
                for key, val in self.attrs.items():
                    f.attrs[key] = val
        else:
            self.flush()
        atexit.register(self.close)

    def output(self, group, **kwargs):
        """Append one record to each dataset of ``group`` (output.py:151-181)."""
        if self.backend == "h5py":
            return self._output_h5(group, **kwargs)
        if group not in self._records:
            self._records[group] = {key: [] for key in kwargs}
        datasets = self._records[group]
        for key in datasets:                          # every dataset must be provided
            if key not in kwargs:
                raise KeyError(key)
        for key, rows in datasets.items():
            val = np.array(_to_host(kwargs[key]))
            if rows and val.shape != rows[0].shape:
                raise ValueError(f"{group}/{key}: record of shape {val.shape}, dataset "
                                 f"holds {rows[0].shape}")
            rows.append(val)
        self._pending += 1
        if self._pending >= self.flush_every:
            self.flush()

    def _output_h5(self, group, **kwargs):
        import h5py
        with h5py.File(self.filename, "a") as f:
            if group not in f:
                f.create_group(group)
                for key, val in kwargs.items():
                    val = np.array(_to_host(val))
                    f[group].create_dataset(key, shape=(0,) + val.shape, dtype=val.dtype,
                                            maxshape=(None,) + val.shape, chunks=True)
            for key in f[group]:
                dset = f[group][key]
                dset.resize(dset.shape[0] + 1, axis=0)
                dset[-1] = _to_host(kwargs.pop(key))

    def flush(self):
        if self.backend == "h5py":
            return
        data = {"attrs/" + k: np.array(v) for k, v in self.attrs.items()}
        for group, datasets in self._records.items():
            for key, rows in datasets.items():
                data[f"{group}/{key}"] = (np.stack(rows) if rows else np.zeros((0,)))
        tmp = self.filename + ".tmp.npz"
        np.savez(tmp, **data)
        os.replace(tmp, self.filename)
        self._pending = 0

    def close(self):
        if self.backend == "npz" and self._pending:
            self.flush()

    @staticmethod
    def read(filename):
        """-> ``{group: {dataset: array}}`` (and ``"attrs"``) of an ``.npz`` output file."""
        out = {}
        with np.load(filename, allow_pickle=False) as data:
            for key in data.files:
                group, _, name = key.rpartition("/")
                out.setdefault(group, {})[name] = data[key]
        return out


def _to_host(val):
    try:
        import torch
        if isinstance(val, torch.Tensor):
            return val.detach().cpu().numpy()
    except ImportError:
        pass
    return val
