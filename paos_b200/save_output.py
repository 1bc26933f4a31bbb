"""Output path: drop-in for ``paos.core.saveOutput`` (reference ``paos/core/saveOutput.py:120-303``).

``save_output`` / ``save_datacube`` write the reference's layout -- an ``info`` group, then one group ``S##`` per saved
surface (inside one group per simulation for a data cube) holding the kept keys, nested dictionaries as sub-groups --
through one of two back ends chosen by what the image offers:

* ``h5py`` present: a real HDF5 file, dataset for dataset what the reference writes;
* ``h5py`` absent (this image): the same tree flattened to ``group/sub/key`` paths in a NumPy ``.npz`` archive next to the
  requested name (``out.h5`` -> ``out.npz``), readable back with :func:`load_output`.

The tree itself (which keys, which conversions, which values are skipped or refused) is built once by :func:`build_tree`
following ``save_recursively_to_hdf5`` (``saveOutput.py:44-82``) and ``save_retval`` (``:120-163``), so both back ends
store the same content.
"""
import datetime
import logging
import os
from copy import deepcopy

import numpy as np

logger = logging.getLogger("paos_b200")

PROGRAM_NAME = "PAOS (paos_b200 device path)"


def have_h5py():
    try:
        import h5py  # noqa: F401

        return True
    except ImportError:
        return False


def remove_keys(dictionary, keys):
    """``saveOutput.py:12-41``."""
    for k in keys:
        dictionary.pop(k, "key not found")


def _leaf(key, data):
    """One value as the reference stores it (``saveOutput.py:63-82``); ``None`` for a value that is skipped."""
    if isinstance(data, (str, int, float, tuple)):
        return np.asarray(data)
    if isinstance(data, np.ndarray):
        return data
    if isinstance(data, list):
        ascii_list = [n.encode("ascii", "ignore") for n in data]
        return np.asarray(ascii_list, dtype="S10").reshape(len(ascii_list), 1)
    if data is None:
        logger.warning("Key %s is None", key)
        return None
    logger.error("Data type for %s not supported", key)
    raise NameError("Data type not supported")


def _flatten(dictionary, prefix, out):
    for key, data in dictionary.items():
        path = f"{prefix}/{key}" if prefix else str(key)
        if isinstance(data, dict):
            _flatten(data, path, out)
            continue
        if isinstance(data, np.ma.MaskedArray):
            data = np.asarray(data.filled(data.fill_value))  # h5py stores the data of a masked array (its buffer)
        leaf = _leaf(key, data)
        if leaf is not None:
            out[path] = leaf


def info_attrs(file_name):
    """The ``info`` group (``saveOutput.py:85-117``)."""
    from . import __version__

    attrs = {"file_name": file_name, "file_time": datetime.datetime.now().isoformat(), "creator": "paos_b200",
             "program_name": PROGRAM_NAME, "program_version": __version__}
    if have_h5py():
        import h5py

        attrs["HDF5_Version"] = h5py.version.hdf5_version
        attrs["h5py_version"] = h5py.version.version
    return attrs


def retval_tree(retval, keys_to_keep, prefix, out):
    """``save_retval`` (``saveOutput.py:120-163``): one ``S##`` group per saved surface.  As in the reference, with
    ``keys_to_keep=None`` the keys of the FIRST surface decide what is kept for every surface."""
    for index in retval.keys():
        item = dict(retval[index])
        if item.get("aperture") is not None:
            item["aperture"] = dict(item["aperture"].__dict__)
        else:
            item["aperture"] = None
        item["ABCDs"] = dict(item["ABCDs"].__dict__)
        item["ABCDt"] = dict(item["ABCDt"].__dict__)
        if keys_to_keep is None:
            keys_to_keep = list(item.keys())
        remove_keys(item, [k for k in list(item.keys()) if k not in keys_to_keep])
        _flatten(item, f"{prefix}/S{index:02d}" if prefix else f"S{index:02d}", out)
    return out


def build_tree(retvals, file_name, group_names=None, keys_to_keep=None):
    """Flat ``{path: array}`` of a whole output file (``group_names=None``: a single simulation)."""
    tree = {}
    _flatten(info_attrs(file_name), "info", tree)
    if group_names is None:
        retval_tree(retvals, keys_to_keep, "", tree)
    else:
        for name, retval in zip(group_names, retvals):
            retval_tree(retval, keys_to_keep, str(name), tree)
    return tree


def _npz_name(file_name):
    root, ext = os.path.splitext(file_name)
    return file_name if ext == ".npz" else root + ".npz"


def _write(tree, file_name, overwrite):
    """Write the flat tree; returns the path actually written."""
    use_h5 = have_h5py() and not file_name.endswith(".npz")
    target = file_name if use_h5 else _npz_name(file_name)
    if overwrite and os.path.isfile(target):
        os.remove(target)
    if use_h5:
        import h5py

        with h5py.File(target, "a") as out:
            for path, arr in tree.items():
                if arr.dtype.kind == "U":
                    out.create_dataset(path, data=str(arr))
                else:
                    out.create_dataset(path, data=arr)
        return target
    if not file_name.endswith(".npz"):
        logger.warning("h5py is not installed: writing %s (same group layout, flattened paths) instead of %s", target, file_name)
    np.savez(target, **{p.replace("/", "|"): a for p, a in tree.items()})
    return target


def save_output(retval, file_name, keys_to_keep=None, overwrite=True):
    """``saveOutput.py:166-218``.  Returns the path written (the reference returns None)."""
    assert isinstance(retval, dict), "parameter retval must be a dict"
    assert isinstance(file_name, str), "parameter file_name must be a string"
    if keys_to_keep is not None:
        assert isinstance(keys_to_keep, list), "parameter keys_to_keep must be a list of strings"
    return _write(build_tree(retval, file_name, None, keys_to_keep), file_name, overwrite)


def save_datacube(retval_list, file_name, group_names, keys_to_keep=None, overwrite=True):
    """``saveOutput.py:221-303``: several simulations (e.g. one per wavelength) in one file, one group each."""
    assert isinstance(retval_list, list), "parameter retval_list must be a list"
    assert isinstance(file_name, str), "parameter file_name must be a string"
    assert isinstance(group_names, list), "parameter group_names must be a list of strings"
    if keys_to_keep is not None:
        assert isinstance(keys_to_keep, list), "parameter keys_to_keep must be a list of strings"
    return _write(build_tree(retval_list, file_name, group_names, keys_to_keep), file_name, overwrite)


def load_output(file_name):
    """Read a file written by :func:`save_output` / :func:`save_datacube` back into nested dictionaries of arrays."""
    flat = {}
    if file_name.endswith(".npz") or not have_h5py():
        with np.load(_npz_name(file_name), allow_pickle=False) as z:
            flat = {k.replace("|", "/"): z[k] for k in z.files}
    else:
        import h5py

        with h5py.File(file_name, "r") as f:
            f.visititems(lambda name, obj: flat.__setitem__(name, obj[()]) if isinstance(obj, h5py.Dataset) else None)
    tree = {}
    for path, arr in flat.items():
        node = tree
        parts = path.split("/")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = arr
    return tree
