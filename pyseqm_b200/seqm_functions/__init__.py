"""Operator-level interface (SURVEY 8(b) level B) under the reference's module names:
hcore.hcore, fock.fock, diag.sym_eig_trunc, SP2.SP2, pack.pack/unpack, energy.*, scf_loop.scf_loop,
anal_grad.scf_analytic_grad -- each a thin host wrapper over one C-ABI entry point."""
