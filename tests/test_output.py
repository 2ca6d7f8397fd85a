"""OutputFile: the append contract of pystella/output.py:151-181 on the npz container."""
import numpy as np
import pytest

import pystella_b200 as ps


def test_output_file_appends(tmp_path):
    name = str(tmp_path / "run")
    out = ps.OutputFile(name=name, runfile=__file__, mpl=1.0, stepper=ps.LowStorageRK54,
                        grid_shape=(32, 32, 32), flush_every=2)
    for i in range(5):
        out.output("energy", t=0.1 * i, a=1. + i, total=np.array([i, 2. * i]))
        if i % 2 == 0:
            out.output("spectra", t=0.1 * i, scalar=np.arange(6.).reshape(2, 3) * i)
    with pytest.raises(KeyError):
        out.output("energy", t=1., a=2.)             # every dataset must be provided
    with pytest.raises(ValueError):
        out.output("energy", t=1., a=2., total=np.zeros(3))
    out.close()
    data = ps.OutputFile.read(name + ".npz")
    assert np.allclose(data["energy"]["t"], 0.1 * np.arange(5))
    assert data["energy"]["total"].shape == (5, 2)
    assert data["spectra"]["scalar"].shape == (3, 2, 3)
    assert np.array_equal(data["spectra"]["scalar"][2], np.arange(6.).reshape(2, 3) * 4)
    assert str(data["attrs"]["stepper"]) == "LowStorageRK54"
    assert "appends" in str(data["attrs"]["runfile"])
    assert tuple(data["attrs"]["grid_shape"]) == (32, 32, 32)
    with pytest.raises(FileExistsError):
        ps.OutputFile(name=name)


def test_output_file_default_name(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    out = ps.OutputFile()
    out.output("g", x=1)
    out.close()
    assert out.filename.startswith("20") and out.filename.endswith(".npz")
    assert ps.OutputFile.read(out.filename)["g"]["x"].tolist() == [1]


def test_mixed_attributes_are_readable(tmp_path):
    """a mixed list / tuple attribute must not become an unreadable object array."""
    import os
    import pystella_b200 as ps
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        out = ps.OutputFile(name="attrs", grid=(4, 4, 4), mixed=[1, "a", 2.5], nested=((1, 2), 3))
        out.output("energy", t=0.0, a=1.0)
        out.close() if hasattr(out, "close") else out.flush()
        data = ps.OutputFile.read(out.filename)
    finally:
        os.chdir(cwd)
    assert "energy" in data
