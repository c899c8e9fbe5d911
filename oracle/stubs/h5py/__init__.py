"""Empty stand-in for h5py (test infrastructure).  The reference imports h5py at module level
(collisionArray.py:11) but only uses it inside CollisionArray.newFromDirectory; this image has no HDF5
library and every shipped .hdf5 is a Git-LFS pointer, so the tests inject collision tensors in memory."""


class File:  # pylint: disable=too-few-public-methods
    def __init__(self, *args, **kwargs):
        raise OSError("h5py is not installed in this image (oracle/stubs/h5py is an empty stand-in)")
