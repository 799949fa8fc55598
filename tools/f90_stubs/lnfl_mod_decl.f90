! Declarations only (no executable code): the public module data of the reference's LNFL_MOD
! (/root/reference/src/lnfl_mod.f90:5-13), so that translated code that says "USE lnfl_mod, ONLY: NBLM, ISO, XNU0, ..."
! (modm.f90:282-283) finds its arrays.  GET_LNFL itself (binary TAPE3 I/O) is not translated: the golden generator fills
! these arrays from a LineStore.  IIM is the allocated second dimension (250000 in the reference; smaller here).
MODULE LNFL_MOD
   PARAMETER (MXMOL=39,IIM=8192,MXBRDMOL=7)
   INTEGER :: NBLM(mxmol),ISO(mxmol,IIM)
   REAL*8 :: XNU0(mxmol,IIM)
   REAL, dimension(mxmol,iim) ::  DELTNU,E,ALPS,ALPF,X,XG,S0,RMOL,SDEP
   integer*4, dimension(mxbrdmol,mxbrdmol,IIM) ::  brd_mol_flg
   REAL, dimension(mxbrdmol,mxbrdmol,IIM) ::  brd_mol_tmp,brd_mol_hw,brd_mol_shft
END MODULE LNFL_MOD
