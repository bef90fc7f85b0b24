!------------------------------------------------------------------------------------------------------------
! gpu_shim.F90 -- drop-in replacement bodies for RXMD's hot-path entry points, bound to librxmd_b200.so.
!
! Build RXMD with this file INSTEAD of the bodies in src/qeq.F90 (QEq), src/pot.F90 (FORCE) and with the
! MODE_MOVE branch of src/comm.F90 (COPYATOMS) routed here; everything else of the Fortran host (main.F90,
! init.F90, param.F90, fileio.F90, cmdline.F90, module.F90) is unchanged.  The subroutine names, argument lists
! and array shapes are the reference's own (src/qeq.F90:2, src/pot.F90:2, src/comm.F90:2), so callers
! (src/main.F90:27-32,75-84; src/cg.F90) compile untouched.
!
! NOTE: this image has no Fortran compiler, so this file is shipped as source and was not compiled here.
! Its marshalling is exercised by the Python harness (rxmd_b200/host/engine.py), which passes the same
! buffers through the same C entry points.  See INTEGRATION.md.
!------------------------------------------------------------------------------------------------------------
module rxg_binding
use iso_c_binding
implicit none

type, bind(C) :: rxg_config
   integer(c_int) :: device, nbuffer, maxneighbs, maxneighbs10, nmincell
   integer(c_int) :: isQEq, NMAXQEq, isPQEq, isEfield, eFieldDir
   real(c_double) :: QEq_tol, Lex_fqs, eFieldStrength
end type

type, bind(C) :: rxg_ff
   integer(c_int) :: nso, nboty, nvaty, ntoty, nhbty, ntable
   real(c_double) :: vpar1, vpar2, cutoff_vpar30, rctap, rctap2, UDR, UDRi
   ! per atom type
   type(c_ptr) :: Val, Valval, Valangle, Vale, mass, plp1, plp2, nlpopt
   type(c_ptr) :: povun2, povun3, povun4, povun5, povun6, povun7, povun8, pval3, pval5, chi, eta
   ! per bond type
   type(c_ptr) :: cBOp1, cBOp3, cBOp5, pbo2h, pbo4h, pbo6h, pbo2, pbo4, pbo6, swtch
   type(c_ptr) :: rc2, pboc1, pboc3, pboc4, pboc5, ovc, v13cor, Desig, Depi, Depipi, pbe1, pbe2, povun1
   ! per valence-angle type
   type(c_ptr) :: theta00, pval1, pval2, pval4, pval6, pval7, pval8, pval9, pval10
   type(c_ptr) :: ppen1, ppen2, ppen3, ppen4, pcoa1, pcoa2, pcoa3, pcoa4
   ! per torsion / hydrogen-bond type
   type(c_ptr) :: ptor1, ptor2, ptor3, ptor4, V1, V2, V3, pcot1, pcot2
   type(c_ptr) :: phb1, phb2, phb3, r0hb
   type(c_ptr) :: inxn2, inxn3, inxn3hb, inxn4
   type(c_ptr) :: TBL_Evdw, TBL_Eclmb, TBL_Eclmb_QEq
   integer(c_int) :: ntype_pqeq
   type(c_ptr) :: isPolarizable, Zpqeq, Kspqeq, inxnpqeq, TBL_Eclmb_pcc, TBL_Eclmb_psc, TBL_Eclmb_pss
end type

type, bind(C) :: rxg_box
   real(c_double) :: HH(9), HHi(9), lata, latb, latc, LBOX(3), OBOX(3), lcsize(3), nblcsize(3)
   integer(c_int) :: cc(3), nbcc(3), nbnmesh, vprocs(3), vID(3), myparity(3), target_node(6), myid, nprocs
   type(c_ptr) :: nbmesh
end type

type(c_ptr), save :: rxg_handle = c_null_ptr
! set .true. by the host right before its main loop (src/main.F90:47) and .false. after it: inside the loop the wrappers below
! tell the library which arrays it already holds (rxg_hint), so pos crosses PCIe once up and once down per step
logical, save :: rxg_loop_hints = .false.
integer(c_int), parameter :: RXG_HINT_ATOMS_ON_DEVICE = 1, RXG_HINT_Q_ON_DEVICE = 2, RXG_HINT_DEFER_POS = 4, &
                             RXG_HINT_CHARGES_STAY = 8

interface
   integer(c_int) function rxg_create(cfg, h) bind(C, name="rxg_create")
      import; type(rxg_config), intent(in) :: cfg; type(c_ptr), intent(out) :: h
   end function
   integer(c_int) function rxg_set_forcefield(h, ff) bind(C, name="rxg_set_forcefield")
      import; type(c_ptr), value :: h; type(rxg_ff), intent(in) :: ff
   end function
   integer(c_int) function rxg_set_box(h, box) bind(C, name="rxg_set_box")
      import; type(c_ptr), value :: h; type(rxg_box), intent(in) :: box
   end function
   integer(c_int) function rxg_comm_unique_id(id) bind(C, name="rxg_comm_unique_id")
      import; character(c_char) :: id(128)
   end function
   integer(c_int) function rxg_comm_init(h, rank, nranks, id) bind(C, name="rxg_comm_init")
      import; type(c_ptr), value :: h; integer(c_int), value :: rank, nranks; character(c_char) :: id(128)
   end function
   integer(c_int) function rxg_qeq(h, natoms, atype, pos, q, qsfp, qsfv, nstep_qeq) bind(C, name="rxg_qeq")
      import; type(c_ptr), value :: h; integer(c_int) :: natoms, nstep_qeq
      real(c_double) :: atype(*), pos(*), q(*), qsfp(*), qsfv(*)
   end function
   integer(c_int) function rxg_pqeq(h, natoms, atype, pos, q, spos, qsfp, qsfv, nstep_qeq) bind(C, name="rxg_pqeq")
      import; type(c_ptr), value :: h; integer(c_int) :: natoms, nstep_qeq
      real(c_double) :: atype(*), pos(*), q(*), spos(*), qsfp(*), qsfv(*)
   end function
   integer(c_int) function rxg_spos_upload(h, natoms, spos) bind(C, name="rxg_spos_upload")
      import; type(c_ptr), value :: h; integer(c_int), value :: natoms; real(c_double) :: spos(*)
   end function
   integer(c_int) function rxg_spos_download(h, natoms, spos) bind(C, name="rxg_spos_download")
      import; type(c_ptr), value :: h; integer(c_int), value :: natoms; real(c_double) :: spos(*)
   end function
   integer(c_int) function rxg_force(h, natoms, atype, pos, f, q, PE, astr) bind(C, name="rxg_force")
      import; type(c_ptr), value :: h; integer(c_int) :: natoms
      real(c_double) :: atype(*), pos(*), f(*), q(*), PE(0:13), astr(6)
   end function
   integer(c_int) function rxg_move(h, natoms, atype, pos, v, q, qs, qt, qsfp, qsfv) bind(C, name="rxg_move")
      import; type(c_ptr), value :: h; integer(c_int) :: natoms
      real(c_double) :: atype(*), pos(*), v(*), q(*), qs(*), qt(*), qsfp(*), qsfv(*)
   end function
   integer(c_int) function rxg_fetch_bonds(h, nbrlist, BO0) bind(C, name="rxg_fetch_bonds")
      import; type(c_ptr), value :: h; integer(c_int) :: nbrlist(*); real(c_double) :: BO0(*)
   end function
   type(c_ptr) function rxg_last_error(h) bind(C, name="rxg_last_error")
      import; type(c_ptr), value :: h
   end function
   integer(c_int) function rxg_hint(h, flags) bind(C, name="rxg_hint")
      import; type(c_ptr), value :: h; integer(c_int), value :: flags
   end function
   integer(c_int) function rxg_it_timer(h, sec) bind(C, name="rxg_it_timer")
      import; type(c_ptr), value :: h; real(c_double) :: sec(30)
   end function
end interface

contains

!--- the reference's error convention: print the message, MPI_FINALIZE, stop (src/main.F90:403-407)
subroutine rxg_check(rc)
   use atoms, only: myid, ierr
   integer(c_int), intent(in) :: rc
   character(kind=c_char), pointer :: msg(:)
   integer :: n
   if (rc == 0) return
   call c_f_pointer(rxg_last_error(rxg_handle), msg, [512])
   n = 1
   do while (n < 512 .and. msg(n) /= c_null_char); n = n + 1; enddo
   write(6,'(a,i4,1x,512a1)') 'ERROR: librxmd_b200 on rank ', myid, msg(1:n-1)
   call MPI_FINALIZE(ierr)
   stop
end subroutine

!--- called once at the end of INITSYSTEM (src/init.F90:288): hand the module globals to the library
subroutine rxg_setup()
   use atoms; use parameters
   type(rxg_config) :: cfg
   type(rxg_ff), target :: ff
   type(rxg_box) :: box
   character(c_char) :: id(128)
   integer :: devcount
   integer(c_int), allocatable, target, save :: ispol_c(:)
   cfg%device = mod(myid, 8)                       ! one rank per GPU of an 8-GPU node
   cfg%nbuffer = NBUFFER; cfg%maxneighbs = MAXNEIGHBS; cfg%maxneighbs10 = MAXNEIGHBS10; cfg%nmincell = NMINCELL
   cfg%isQEq = isQEq; cfg%NMAXQEq = NMAXQEq; cfg%isPQEq = merge(1, 0, isPQEq); cfg%isEfield = merge(1, 0, isEfield)
   cfg%eFieldDir = eFieldDir; cfg%QEq_tol = QEq_tol; cfg%Lex_fqs = Lex_fqs; cfg%eFieldStrength = eFieldStrength
   call rxg_check(rxg_create(cfg, rxg_handle))
   ff%nso = nso; ff%nboty = nboty; ff%nvaty = size(theta00); ff%ntoty = size(V1); ff%nhbty = size(r0hb); ff%ntable = NTABLE
   ff%vpar1 = vpar1; ff%vpar2 = vpar2; ff%cutoff_vpar30 = cutoff_vpar30
   ff%rctap = rctap; ff%rctap2 = rctap2; ff%UDR = UDR; ff%UDRi = UDRi
   ff%Val = c_loc(Val); ff%Valval = c_loc(Valval); ff%Valangle = c_loc(Valangle); ff%Vale = c_loc(Vale); ff%mass = c_loc(mass)
   ff%plp1 = c_loc(plp1); ff%plp2 = c_loc(plp2); ff%nlpopt = c_loc(nlpopt)
   ff%povun2 = c_loc(povun2); ff%povun3 = c_loc(povun3); ff%povun4 = c_loc(povun4); ff%povun5 = c_loc(povun5)
   ff%povun6 = c_loc(povun6); ff%povun7 = c_loc(povun7); ff%povun8 = c_loc(povun8)
   ff%pval3 = c_loc(pval3); ff%pval5 = c_loc(pval5); ff%chi = c_loc(chi); ff%eta = c_loc(eta)
   ff%cBOp1 = c_loc(cBOp1); ff%cBOp3 = c_loc(cBOp3); ff%cBOp5 = c_loc(cBOp5)
   ff%pbo2h = c_loc(pbo2h); ff%pbo4h = c_loc(pbo4h); ff%pbo6h = c_loc(pbo6h)
   ff%pbo2 = c_loc(pbo2); ff%pbo4 = c_loc(pbo4); ff%pbo6 = c_loc(pbo6); ff%swtch = c_loc(switch)
   ff%rc2 = c_loc(rc2); ff%pboc1 = c_loc(pboc1); ff%pboc3 = c_loc(pboc3); ff%pboc4 = c_loc(pboc4); ff%pboc5 = c_loc(pboc5)
   ff%ovc = c_loc(ovc); ff%v13cor = c_loc(v13cor); ff%Desig = c_loc(Desig); ff%Depi = c_loc(Depi); ff%Depipi = c_loc(Depipi)
   ff%pbe1 = c_loc(pbe1); ff%pbe2 = c_loc(pbe2); ff%povun1 = c_loc(povun1)
   ff%theta00 = c_loc(theta00); ff%pval1 = c_loc(pval1); ff%pval2 = c_loc(pval2); ff%pval4 = c_loc(pval4)
   ff%pval6 = c_loc(pval6); ff%pval7 = c_loc(pval7); ff%pval8 = c_loc(pval8); ff%pval9 = c_loc(pval9); ff%pval10 = c_loc(pval10)
   ff%ppen1 = c_loc(ppen1); ff%ppen2 = c_loc(ppen2); ff%ppen3 = c_loc(ppen3); ff%ppen4 = c_loc(ppen4)
   ff%pcoa1 = c_loc(pcoa1); ff%pcoa2 = c_loc(pcoa2); ff%pcoa3 = c_loc(pcoa3); ff%pcoa4 = c_loc(pcoa4)
   ff%ptor1 = c_loc(ptor1); ff%ptor2 = c_loc(ptor2); ff%ptor3 = c_loc(ptor3); ff%ptor4 = c_loc(ptor4)
   ff%V1 = c_loc(V1); ff%V2 = c_loc(V2); ff%V3 = c_loc(V3); ff%pcot1 = c_loc(pcot1); ff%pcot2 = c_loc(pcot2)
   ff%phb1 = c_loc(phb1); ff%phb2 = c_loc(phb2); ff%phb3 = c_loc(phb3); ff%r0hb = c_loc(r0hb)
   ff%inxn2 = c_loc(inxn2); ff%inxn3 = c_loc(inxn3); ff%inxn3hb = c_loc(inxn3hb); ff%inxn4 = c_loc(inxn4)
   ff%TBL_Evdw = c_loc(TBL_Evdw); ff%TBL_Eclmb = c_loc(TBL_Eclmb); ff%TBL_Eclmb_QEq = c_loc(TBL_Eclmb_QEq)
   ff%ntype_pqeq = 0
   if (isPQEq) then                                ! module pqeq_vars after initialize_pqeq (src/module.F90:488-613)
      ff%ntype_pqeq = ntype_pqeq
      allocate(ispol_c(ntype_pqeq)); ispol_c = merge(1_c_int, 0_c_int, isPolarizable)     ! logical -> int, kept alive
      ff%isPolarizable = c_loc(ispol_c); ff%Zpqeq = c_loc(Zpqeq); ff%Kspqeq = c_loc(Kspqeq); ff%inxnpqeq = c_loc(inxnpqeq)
      ff%TBL_Eclmb_pcc = c_loc(TBL_Eclmb_pcc); ff%TBL_Eclmb_psc = c_loc(TBL_Eclmb_psc); ff%TBL_Eclmb_pss = c_loc(TBL_Eclmb_pss)
   endif
   call rxg_check(rxg_set_forcefield(rxg_handle, ff))
   box%HH = reshape(HH(:,:,0), [9]); box%HHi = reshape(HHi, [9])
   box%lata = lata; box%latb = latb; box%latc = latc
   box%LBOX = LBOX(1:3); box%OBOX = OBOX; box%lcsize = lcsize; box%nblcsize = nblcsize
   box%cc = cc; box%nbcc = nbcc; box%nbnmesh = nbnmesh; box%vprocs = vprocs; box%vID = vID; box%myparity = myparity
   box%target_node = target_node; box%myid = myid; box%nprocs = nprocs; box%nbmesh = c_loc(nbmesh)
   call rxg_check(rxg_set_box(rxg_handle, box))
   if (nprocs > 1) then
      if (myid == 0) call rxg_check(rxg_comm_unique_id(id))
      call MPI_BCAST(id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
      call rxg_check(rxg_comm_init(rxg_handle, myid, nprocs, id))
   endif
end subroutine

end module rxg_binding

!------------------------------------------------------------------------------------------------------------
subroutine QEq(atype, pos, q)                                  ! replaces src/qeq.F90:2-178
use atoms; use rxg_binding
implicit none
real(8) :: atype(NBUFFER), pos(NBUFFER,3), q(NBUFFER)
! inside the main loop nothing touches atype/pos/q between COPYATOMS(MODE_MOVE), QEq and FORCE (src/main.F90:75-84)
if (rxg_loop_hints) call rxg_check(rxg_hint(rxg_handle, RXG_HINT_ATOMS_ON_DEVICE + RXG_HINT_Q_ON_DEVICE + RXG_HINT_DEFER_POS))
call rxg_check(rxg_qeq(rxg_handle, NATOMS, atype, pos, q, qsfp, qsfv, nstep_qeq))
end subroutine

!------------------------------------------------------------------------------------------------------------
subroutine PQEq(atype, pos, q)                                 ! replaces src/pqeq.F90:2-182 (spos is module state, relaxed at the end)
use atoms; use rxg_binding
implicit none
real(8) :: atype(NBUFFER), pos(NBUFFER,3), q(NBUFFER)
if (rxg_loop_hints) call rxg_check(rxg_hint(rxg_handle, RXG_HINT_ATOMS_ON_DEVICE + RXG_HINT_Q_ON_DEVICE + RXG_HINT_DEFER_POS))
call rxg_check(rxg_pqeq(rxg_handle, NATOMS, atype, pos, q, spos, qsfp, qsfv, nstep_qeq))
end subroutine

!------------------------------------------------------------------------------------------------------------
subroutine FORCE(atype, pos, f, q)                             ! replaces src/pot.F90:2-90
use atoms; use rxg_binding
implicit none
real(8) :: atype(NBUFFER), q(NBUFFER), pos(NBUFFER,3), f(NBUFFER,3)
if (rxg_loop_hints) call rxg_check(rxg_hint(rxg_handle, RXG_HINT_ATOMS_ON_DEVICE + RXG_HINT_Q_ON_DEVICE))
call rxg_check(rxg_force(rxg_handle, NATOMS, atype, pos, f, q, PE, astr))      ! hands back the final pos of the step
end subroutine

!------------------------------------------------------------------------------------------------------------
! it_timer(1:30) for the timing table of src/main.F90:135-180: the library measures its phases with CUDA events in the
! reference's own slots (seconds); the host keeps ticks, so convert with its clock rate irt.  Call once before the table.
subroutine rxg_fill_it_timer()
use atoms; use rxg_binding
implicit none
real(c_double) :: sec(30)
integer :: k
call rxg_check(rxg_it_timer(rxg_handle, sec))
do k = 1, 19
   if (k /= 2 .and. k /= 14 .and. k /= 17) it_timer(k) = it_timer(k) + nint(sec(k) * irt)
enddo
it_timer(24) = it_timer(24) + nint(sec(24))       ! QEq iterations (src/qeq.F90:172)
end subroutine

!------------------------------------------------------------------------------------------------------------
subroutine COPYATOMS(imode, dr, atype, pos, v, f, q)           ! replaces src/comm.F90:2-100 for the host's MODE_MOVE call
use atoms; use rxg_binding
implicit none
integer, intent(in) :: imode
real(8), intent(in) :: dr(3)
real(8) :: atype(NBUFFER), q(NBUFFER), pos(NBUFFER,3), v(NBUFFER,3), f(NBUFFER,3)
integer(c_int) :: ihint
if (imode /= MODE_MOVE) then
   print'(a,i3)', "ERROR: imode doesn't match in COPYATOMS: ", imode    ! the other modes run inside the library
   call MPI_FINALIZE(ierr); stop
endif
if (isPQEq) call rxg_check(rxg_spos_upload(rxg_handle, NATOMS, spos))      ! spos migrates with the atom (src/comm.F90:153,165-167)
! in a step without output the host does not read pos before FORCE returns it: leave the ulp-level round trip on the device
! inside the main loop nothing reads or writes q, qs, qt between the QEq of one step and the QEq of the next except the
! charge-Lagrangian lines, which read q AFTER this call's QEq returned it (src/main.F90:64-98): they migrate on the device
! (only in steps whose QEq runs, mod(nstep,qstep)==0: otherwise the host's q must follow the migration)
ihint = 0
if (rxg_loop_hints .and. mod(nstep, fstep) /= 0) ihint = ihint + RXG_HINT_DEFER_POS
if (rxg_loop_hints .and. mod(nstep, qstep) == 0) ihint = ihint + RXG_HINT_CHARGES_STAY
if (ihint /= 0) call rxg_check(rxg_hint(rxg_handle, ihint))
call rxg_check(rxg_move(rxg_handle, NATOMS, atype, pos, v, q, qs, qt, qsfp, qsfv))
if (isPQEq) call rxg_check(rxg_spos_download(rxg_handle, NATOMS, spos))
end subroutine
