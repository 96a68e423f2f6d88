!
!  mcpolar_patch.f90 -- what a maintainer adds to the reference's src/mcpolar.f90 to run the photon
!  loop on B200s through libtamc.so.  Not a stand-alone program: three fragments, each marked with the
!  reference lines it goes next to or replaces.  Everything else in mcpolar.f90 (input.params, gridset,
!  init_opt1, the 3dFD heat / ablation coupling, writer) is unchanged.  Shipped as source; there is no
!  Fortran compiler in this image (INTEGRATION.md).
!
!  One MPI rank drives one GPU (rank id -> device mod(id, ngpus)), exactly as the reference gives each
!  rank its own `do j = 1, nphotons` loop; the ranks' tallies are summed by NCCL over NVLink instead of
!  MPI_allREDUCE.
!

!--- (1) declarations, next to mcpolar.f90:26-34 ---------------------------------------------------
      use tamc_mod
      type(c_ptr)            :: tamc
      integer(c_int)         :: ierr, oflags
      character(kind=c_char) :: nccl_id(128)
      integer(c_int64_t)     :: seed64

!--- (2) set-up, after gridset/delta at mcpolar.f90:109-112 ----------------------------------------
      ierr = tamc_init(int(mod(id, tamc_device_count()), c_int), nxg, nyg, nzg, xmax, ymax, zmax, delta, tamc)
      call tamc_check(ierr, 'tamc_init')
      !  every rank needs the same NCCL id: rank 0 makes it, MPI carries it once
      if(id == 0)then
         ierr = tamc_comm_unique_id(nccl_id)
         call tamc_check(ierr, 'tamc_comm_unique_id')
      end if
      call MPI_Bcast(nccl_id, 128, MPI_CHARACTER, 0, new_comm)
      !  optional, BEFORE tamc_comm_init and on every rank: only rank 0's rhokap is read (it reaches the other GPUs over
      !  NVLink) and only rank 0's jmeanGLOBAL is written -- all the reference consumes (3dFD.f90:95 scatters rank 0's copy;
      !  setupThermalCoeff, :312-361, leaves identical rhokap on every rank).  Every rank still makes the same call.
      ! ierr = tamc_set_option(tamc, 'root_io'//c_null_char, 1_c_int64_t)
      ierr = tamc_comm_init(tamc, int(numproc, c_int), int(id, c_int), nccl_id)
      call tamc_check(ierr, 'tamc_comm_init')
      !  optional: page-lock the two arrays that cross PCIe every MC call
      ierr = tamc_pin_host(rhokap, int(size(rhokap), c_int64_t)*8_c_int64_t)
      call tamc_check(ierr, 'tamc_pin_host(rhokap)')
      ierr = tamc_pin_host(jmeanGLOBAL, int(size(jmeanGLOBAL), c_int64_t)*8_c_int64_t)
      call tamc_check(ierr, 'tamc_pin_host(jmeanGLOBAL)')
      seed64 = 95648324_c_int64_t          ! mcpolar.f90:97: the run seed; ranks are told apart by packet id
      oflags = 0                           ! TAMC_SCATTER to run the albedo/stokes loop instead of the stub

!--- (3) the hot path: REPLACES mcpolar.f90:151-173 (photon loop + MPI_allREDUCE) --------------------
      !  rhokap was rewritten by setupThermalCoeff at the end of the previous iteration (mcpolar.f90:182):
      !  upload it, run nphotons packets on this rank, sum the tally over all ranks, UNSCALED sum into
      !  jmeanGLOBAL.  One call, so the library can overlap the two PCIe copies with the transport
      !  (equivalent to tamc_set_optics followed by tamc_run, which remain available).
      ierr = tamc_run_optics(tamc, rhokap, albedo, hgg, n1, n2, oflags, int(nphotons, c_int64_t), seed64, &
                             jmeanGLOBAL, c_null_ptr)
      call tamc_check(ierr, 'tamc_run_optics')
      !  mcpolar.f90:174 (the power / packet-count / voxel-volume scaling) stays exactly as it is.

!--- (4) before MPI_Finalize at mcpolar.f90:215 ------------------------------------------------------
      ierr = tamc_finalize(tamc)
