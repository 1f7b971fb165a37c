"""Registers the CUDA engine in the REFERENCE's own facade, at run time.

``install()`` does to an imported, unmodified ``pyconmech`` exactly what INTEGRATION.md section 1
asks a maintainer to write into ``src/pyconmech/frame_analysis/stiffness_checker.py:57-66``: a third
branch ``checker_engine == "cuda"`` that constructs :class:`conmech_b200.CudaStiffness`, and the
batched ``solve_many``.  Every other method of the reference facade is used as it is - it only ever
talks to ``self._sc_ins`` through the ``StiffnessBase`` interface (``stiffness_base.py:8-155``), and the
reference's ``Model`` / ``LoadCase`` objects are consumed directly (same attribute names as
``conmech_b200.io_base``).

    import pyconmech
    from conmech_b200 import pyconmech_adapter
    pyconmech_adapter.install()
    sc = pyconmech.StiffnessChecker.from_json(path, checker_engine="cuda")

tests/test_reference_suite.py runs the reference's own pytest suite through this switch.
"""


def install(device=None):
    import pyconmech.frame_analysis.stiffness_checker as ref
    from .cuda_stiffness import CudaStiffness
    cls = ref.StiffnessChecker
    if getattr(cls, '_conmech_b200_installed', False):
        return cls
    numpy_init = cls.__init__

    def __init__(self, model, verbose=False, checker_engine=ref.DEFAULT_ENGINE):
        if checker_engine != 'cuda':
            return numpy_init(self, model, verbose=verbose, checker_engine=checker_engine)
        # the same statements the reference runs for its own engines (stiffness_checker.py:66-76)
        self._sc_ins = CudaStiffness(model, verbose=verbose, device=device)
        self.set_nodal_displacement_tol()
        self.model = model
        if not len(model.supports) > 0:
            raise RuntimeError('there needs to be at least one support (fixed) vertex in the model!')

    def solve_many(self, list_of_existing_ids, return_details=False):
        r = self._sc_ins.solve_many(list_of_existing_ids, want=('flag', 'status', 'max_trans', 'compliance'))
        flags = [bool(f) for f in r['flag'][:, 0]]
        if return_details:
            return flags, {k: r[k][:, 0].copy() for k in ('status', 'max_trans', 'compliance')}
        return flags

    cls.__init__ = __init__
    cls.solve_many = solve_many
    cls._conmech_b200_installed = True
    return cls
