"""Import alias: the product package lives in ``aliasfree-diffusion-models-pytorch_b200/``
(the directory name the build contract asks for, which is not a valid Python identifier).
``import aliasfree_b200`` executes that package under this name."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "aliasfree-diffusion-models-pytorch_b200")
__path__ = [_pkg_dir]
__file__ = _os.path.join(_pkg_dir, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
