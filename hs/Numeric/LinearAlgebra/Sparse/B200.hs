{-# LANGUAGE ForeignFunctionInterface #-}
-- |
-- Module      : Numeric.LinearAlgebra.Sparse.B200
-- Description : FFI shim that puts libsla_b200.so behind the operator surface of
--               Numeric.LinearAlgebra.Sparse (ocramz/sparse-linear-algebra).
--
-- NOT COMPILED in the build image (no GHC there).  It shows the binding a maintainer would add:
-- the names, argument order and exceptions are those of the reference
-- (src/Numeric/LinearAlgebra/Class.hs:57-99, 126-153, 195-229; src/Numeric/LinearAlgebra/Sparse.hs:630-667,
-- 855-981, 1016-1072).  `bicgstabStep` etc. are monomorphic in the container in the reference
-- (BICGSTAB a holds SpVector a, Sparse.hs:962-963), so the drop-in is this module exporting the same names
-- over opaque device handles plus `toDevice` / `fromDevice` marshalling, not a new class instance.
module Numeric.LinearAlgebra.Sparse.B200
  ( Ctx, DMatrix, DVector, DDense, BICGSTAB(..), CGS(..), CGNE(..)
  , withB200, toDeviceSM, toDeviceSV, fromDeviceSV
  , (#>), (<#), (<.>), (^+^), (^-^), (.*), (./), norm2, normalize2, transpose
  , (##), (##^), (#^#), normFrobenius, toDeviceDense, fromDeviceDense
  , bicgsInit, bicgstabStep, bicgstabStepInPlace, cgsInit, cgsStep, cgsStepInPlace, cgneInit, cgneStep, cgneStepInPlace
  , LinSolveMethod(..), linSolve0, arnoldi, basisColumn, basisToHost, (<\>)
  , diagPartitions, jacobiPre, mSsorPre, ilu0Pre, triLowerSolve, triUpperSolve
    -- * one process, several GPUs (sla_init_multi): global row-partitioned objects, the caller never sees ranks
  , MCtx, MMatrix, MVector, withB200Multi, toDeviceSMMulti, toDeviceSVMulti, fromDeviceSVMulti, matVecMulti, dotMulti, linSolve0Multi, gmresMulti
  ) where

import Control.Exception (bracket, throwIO)
import Control.Monad (when)
import Data.Int (Int32, Int64)
import qualified Data.Vector.Storable as VS
import Foreign
import Foreign.C.String (CString, peekCString)
import Foreign.C.Types

-- the reference's own types, used only for marshalling and for the exceptions we re-throw
import Control.Exception.Common (OperandSizeMismatch(..), IterationException(..), MatrixException(..))
import Data.Sparse.SpMatrix (SpMatrix, immSM, nrows, ncols)
import Data.Sparse.SpVector (SpVector, fromListDenseSV, toDenseListSV, dim)
import qualified Data.Sparse.Internal.IntM as I
import Data.Foldable (toList)

data SlaCtx; data SlaCsr; data SlaVec; data SlaKrylov; data SlaDense
newtype Ctx     = Ctx (Ptr SlaCtx)
data DMatrix    = DMatrix Ctx (ForeignPtr SlaCsr)
data DVector    = DVector Ctx (ForeignPtr SlaVec)
-- | dense block on the device, released by sla_dense_free when the Haskell value dies: the Arnoldi basis Q (column-major
--   fp64) or a row-major dense operand / result of (##)
data DDense     = DDense Ctx (ForeignPtr SlaDense)

type Status = CInt

foreign import ccall safe "sla_init"            c_init        :: CInt -> Ptr (Ptr SlaCtx) -> IO Status
foreign import ccall safe "sla_finalize"        c_finalize    :: Ptr SlaCtx -> IO ()
foreign import ccall safe "sla_last_error"      c_last_error  :: Ptr SlaCtx -> IO CString
foreign import ccall safe "sla_csr_from_coo"    c_from_coo    :: Ptr SlaCtx -> Int64 -> Int64 -> Int64 -> Ptr Int64 -> Ptr Int64 -> Ptr Double -> Ptr (Ptr SlaCsr) -> IO Status
foreign import ccall safe "sla_csr_dims"        c_csr_dims    :: Ptr SlaCsr -> Ptr Int64 -> Ptr Int64 -> Ptr Int64 -> IO Status
foreign import ccall safe "sla_csr_transpose"   c_transpose   :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr (Ptr SlaCsr) -> IO Status
foreign import ccall safe "&sla_csr_free"       p_csr_free    :: FunPtr (Ptr SlaCsr -> IO ())
foreign import ccall safe "sla_vec_from_host"   c_vec_from    :: Ptr SlaCtx -> Int64 -> Ptr Double -> Ptr (Ptr SlaVec) -> IO Status
foreign import ccall safe "sla_vec_create"      c_vec_create  :: Ptr SlaCtx -> Int64 -> Ptr (Ptr SlaVec) -> IO Status
foreign import ccall safe "sla_vec_to_host"     c_vec_to      :: Ptr SlaCtx -> Ptr SlaVec -> Ptr Double -> IO Status
foreign import ccall safe "sla_vec_dim"         c_vec_dim     :: Ptr SlaVec -> IO Int64
foreign import ccall safe "&sla_vec_free"       p_vec_free    :: FunPtr (Ptr SlaVec -> IO ())
foreign import ccall safe "sla_spmv"            c_spmv        :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> IO Status
foreign import ccall safe "sla_spmvT"           c_spmvT       :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> IO Status
foreign import ccall safe "sla_dot"             c_dot         :: Ptr SlaCtx -> Ptr SlaVec -> Ptr SlaVec -> Ptr Double -> IO Status
foreign import ccall safe "sla_norm2"           c_norm2       :: Ptr SlaCtx -> Ptr SlaVec -> Ptr Double -> IO Status
foreign import ccall safe "sla_vec_add"         c_add         :: Ptr SlaCtx -> Ptr SlaVec -> Ptr SlaVec -> Ptr SlaVec -> IO Status
foreign import ccall safe "sla_vec_sub"         c_sub         :: Ptr SlaCtx -> Ptr SlaVec -> Ptr SlaVec -> Ptr SlaVec -> IO Status
foreign import ccall safe "sla_vec_scale"       c_scale       :: Ptr SlaCtx -> Double -> Ptr SlaVec -> Ptr SlaVec -> IO Status
foreign import ccall safe "sla_vec_normalize2"  c_normalize2  :: Ptr SlaCtx -> Ptr SlaVec -> Ptr SlaVec -> IO Status
foreign import ccall safe "sla_bicgstab_init"   c_bicg_init   :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> Ptr (Ptr SlaKrylov) -> IO Status
foreign import ccall safe "sla_bicgstab_step"   c_bicg_step   :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaKrylov -> IO Status
foreign import ccall safe "sla_cgs_init"        c_cgs_init    :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> Ptr (Ptr SlaKrylov) -> IO Status
foreign import ccall safe "sla_cgs_step"        c_cgs_step    :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaKrylov -> IO Status
foreign import ccall safe "sla_cgne_init"       c_cgne_init   :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> Ptr (Ptr SlaKrylov) -> IO Status
foreign import ccall safe "sla_cgne_step"       c_cgne_step   :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaKrylov -> IO Status
foreign import ccall safe "sla_csr_diag_partitions" c_diag_parts :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr (Ptr SlaCsr) -> Ptr (Ptr SlaCsr) -> Ptr (Ptr SlaCsr) -> IO Status
foreign import ccall safe "sla_jacobi_pre"      c_jacobi_pre  :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr (Ptr SlaCsr) -> IO Status
foreign import ccall safe "sla_mssor_pre"       c_mssor_pre   :: Ptr SlaCtx -> Ptr SlaCsr -> Double -> Ptr (Ptr SlaCsr) -> Ptr (Ptr SlaCsr) -> IO Status
foreign import ccall safe "sla_tri_lower_solve" c_tri_lower   :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> IO Status
foreign import ccall safe "sla_tri_upper_solve" c_tri_upper   :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> IO Status
foreign import ccall safe "sla_krylov_view"     c_kry_view    :: Ptr SlaCtx -> Ptr SlaKrylov -> CInt -> Ptr (Ptr SlaVec) -> IO Status
foreign import ccall safe "&sla_krylov_free"    p_kry_free    :: FunPtr (Ptr SlaKrylov -> IO ())
foreign import ccall safe "sla_linsolve0"       c_linsolve0   :: Ptr SlaCtx -> CInt -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> Ptr () -> Ptr SlaVec -> Ptr CInt -> Ptr Double -> IO Status
foreign import ccall safe "sla_gmres"           c_gmres       :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> CInt -> Ptr () -> Ptr SlaVec -> Ptr CInt -> Ptr Double -> IO Status
foreign import ccall safe "sla_arnoldi"         c_arnoldi     :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> CInt -> Ptr (Ptr SlaDense) -> Ptr Double -> Ptr CInt -> IO Status
foreign import ccall safe "sla_krylov_clone"    c_kry_clone   :: Ptr SlaCtx -> Ptr SlaKrylov -> Ptr (Ptr SlaKrylov) -> IO Status
foreign import ccall safe "&sla_dense_free"     p_dense_free  :: FunPtr (Ptr SlaDense -> IO ())
foreign import ccall safe "sla_dense_dims"      c_dense_dims  :: Ptr SlaDense -> Ptr Int64 -> Ptr Int64 -> IO Status
foreign import ccall safe "sla_dense_to_host"   c_dense_to    :: Ptr SlaCtx -> Ptr SlaDense -> Ptr Double -> IO Status          -- column-major blocks (Q)
foreign import ccall safe "sla_dense_to_host_f64" c_dense_to_rm :: Ptr SlaCtx -> Ptr SlaDense -> Ptr Double -> IO Status        -- row-major blocks
foreign import ccall safe "sla_dense_column"    c_dense_col   :: Ptr SlaCtx -> Ptr SlaDense -> Int64 -> Ptr (Ptr SlaVec) -> IO Status
foreign import ccall safe "sla_dense_create"    c_dense_new   :: Ptr SlaCtx -> Int64 -> Int64 -> CInt -> Ptr (Ptr SlaDense) -> IO Status
foreign import ccall safe "sla_dense_from_host" c_dense_from  :: Ptr SlaCtx -> Int64 -> Int64 -> Ptr Double -> CInt -> Ptr (Ptr SlaDense) -> IO Status
foreign import ccall safe "sla_spmm_dense"      c_spmm        :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaDense -> Ptr SlaDense -> IO Status
foreign import ccall safe "sla_spmm_dense_abt"  c_spmm_abt    :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaDense -> Ptr SlaDense -> IO Status
foreign import ccall safe "sla_spmm_dense_atb"  c_spmm_atb    :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaDense -> Ptr SlaDense -> IO Status
foreign import ccall safe "sla_csr_norm_frobenius" c_frob     :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr Double -> IO Status
foreign import ccall safe "sla_ilu0_pre"        c_ilu0_pre    :: Ptr SlaCtx -> Ptr SlaCsr -> Ptr (Ptr SlaCsr) -> Ptr (Ptr SlaCsr) -> IO Status

-- | status code -> the reference's exception (Control/Exception/Common.hs:44-76)
check :: Ctx -> String -> Status -> IO ()
check (Ctx c) who st = when (st /= 0) $ do
  msg <- c_last_error c >>= peekCString
  case st of
    1 -> throwIO (MatVecSizeMismatchException who (0, 0) 0)             -- SLA_ERR_SIZE_MISMATCH
    2 -> ioError (userError "insertSpMatrix : index out of bounds")     -- SLA_ERR_OOB_INDEX  (SpMatrix.hs:205-208)
    3 -> throwIO (IterE who msg :: IterationException ())               -- SLA_ERR_UNSUPPORTED_METHOD
    10 -> throwIO (NeedsPivoting who msg :: MatrixException Double)     -- SLA_ERR_NEEDS_PIVOTING (Sparse.hs:757, 791)
    _ -> ioError (userError (who ++ ": " ++ msg))

withB200 :: Int -> (Ctx -> IO a) -> IO a
withB200 dev = bracket open (\(Ctx c) -> c_finalize c)
  where open = alloca $ \pp -> do { st <- c_init (fromIntegral dev) pp; c <- peek pp
                                  ; when (st /= 0) (ioError (userError "sla_init failed (no CUDA device: there is no CPU path)"))
                                  ; return (Ctx c) }

-- | Marshal an SpMatrix: walk immSM in ASCENDING (row, col) order — do NOT use toListSM, it conses and returns
--   descending order (SpMatrix.hs:251-253).  The library sorts and de-duplicates anyway (last write wins).
toDeviceSM :: Ctx -> SpMatrix Double -> IO DMatrix
toDeviceSM ctx@(Ctx c) sm = do
  let trip = [ (i, j, x) | (i, row) <- I.toList (immSM sm), (j, x) <- I.toList row ]
      is = VS.fromList [ fromIntegral i | (i, _, _) <- trip ] :: VS.Vector Int64
      js = VS.fromList [ fromIntegral j | (_, j, _) <- trip ] :: VS.Vector Int64
      vs = VS.fromList [ x | (_, _, x) <- trip ]
  VS.unsafeWith is $ \pi' -> VS.unsafeWith js $ \pj -> VS.unsafeWith vs $ \pv -> alloca $ \pp -> do
    c_from_coo c (fromIntegral (nrows sm)) (fromIntegral (ncols sm)) (fromIntegral (VS.length vs)) pi' pj pv pp >>= check ctx "fromListSM"
    h <- peek pp
    DMatrix ctx <$> newForeignPtr p_csr_free h

-- | Absent keys marshal as 0.0 (toDenseListSV, SpVector.hs:300-301).
toDeviceSV :: Ctx -> SpVector Double -> IO DVector
toDeviceSV ctx@(Ctx c) v = VS.unsafeWith (VS.fromList (toDenseListSV v)) $ \px -> alloca $ \pp -> do
  c_vec_from c (fromIntegral (dim v)) px pp >>= check ctx "toDeviceSV"
  peek pp >>= fmap (DVector ctx) . newForeignPtr p_vec_free

fromDeviceSV :: DVector -> IO (SpVector Double)
fromDeviceSV (DVector ctx@(Ctx c) fv) = withForeignPtr fv $ \pv -> do
  n <- fromIntegral <$> c_vec_dim pv
  allocaArray n $ \px -> do
    c_vec_to c pv px >>= check ctx "fromDeviceSV"
    fromListDenseSV n <$> peekArray n px

newVec :: Ctx -> Int64 -> IO DVector
newVec ctx@(Ctx c) n = alloca $ \pp -> do
  c_vec_create c n pp >>= check ctx "zeroSV"
  peek pp >>= fmap (DVector ctx) . newForeignPtr p_vec_free

dimD :: DVector -> IO Int64
dimD (DVector _ fv) = withForeignPtr fv c_vec_dim

binop :: String -> (Ptr SlaCtx -> Ptr SlaVec -> Ptr SlaVec -> Ptr SlaVec -> IO Status) -> DVector -> DVector -> IO DVector
binop who f x@(DVector ctx@(Ctx c) fx) (DVector _ fy) = do
  z@(DVector _ fz) <- dimD x >>= newVec ctx
  withForeignPtr fx $ \px -> withForeignPtr fy $ \py -> withForeignPtr fz $ \pz -> f c px py pz >>= check ctx who
  return z

infixl 6 ^+^, ^-^
infixr 7 .*, ./
(^+^), (^-^) :: DVector -> DVector -> IO DVector
(^+^) = binop "^+^" c_add
(^-^) = binop "^-^" c_sub

(.*) :: Double -> DVector -> IO DVector
a .* x@(DVector ctx@(Ctx c) fx) = do
  z@(DVector _ fz) <- dimD x >>= newVec ctx
  withForeignPtr fx $ \px -> withForeignPtr fz $ \pz -> c_scale c a px pz >>= check ctx ".*"
  return z

(./) :: DVector -> Double -> IO DVector
v ./ s = recip s .* v                                   -- Class.hs:94-95

(<.>) :: DVector -> DVector -> IO Double
(DVector ctx@(Ctx c) fx) <.> (DVector _ fy) =
  withForeignPtr fx $ \px -> withForeignPtr fy $ \py -> alloca $ \po -> c_dot c px py po >>= check ctx "<.>" >> peek po

norm2 :: DVector -> IO Double
norm2 (DVector ctx@(Ctx c) fx) = withForeignPtr fx $ \px -> alloca $ \po -> c_norm2 c px po >>= check ctx "norm2" >> peek po

normalize2 :: DVector -> IO DVector
normalize2 x@(DVector ctx@(Ctx c) fx) = do
  z@(DVector _ fz) <- dimD x >>= newVec ctx
  withForeignPtr fx $ \px -> withForeignPtr fz $ \pz -> c_normalize2 c px pz >>= check ctx "normalize2"
  return z

-- | (nrows, ncols) of a device matrix
dimM :: DMatrix -> IO (Int64, Int64)
dimM (DMatrix _ fa) = withForeignPtr fa $ \pa -> alloca $ \pm -> alloca $ \pn -> alloca $ \pz -> do
  _ <- c_csr_dims pa pm pn pz
  (,) <$> peek pm <*> peek pn

matvecWith :: String -> (Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> IO Status) -> Int64 -> DMatrix -> DVector -> IO DVector
matvecWith who f n (DMatrix ctx@(Ctx c) fa) (DVector _ fx) = do
  y@(DVector _ fy) <- newVec ctx n
  withForeignPtr fa $ \pa -> withForeignPtr fx $ \px -> withForeignPtr fy $ \py -> f c pa px py >>= check ctx who
  return y

-- | aa #> v   (Common.hs:242-250)          v <# aa   (Common.hs:253-256)
(#>) :: DMatrix -> DVector -> IO DVector
aa #> v = do { (m, _) <- dimM aa; matvecWith "matVec" c_spmv m aa v }
(<#) :: DVector -> DMatrix -> IO DVector
v <# aa = do { (_, n) <- dimM aa; matvecWith "vecMat" c_spmvT n aa v }

transpose :: DMatrix -> IO DMatrix
transpose (DMatrix ctx@(Ctx c) fa) = withForeignPtr fa $ \pa -> alloca $ \pp -> do
  c_transpose c pa pp >>= check ctx "transpose"
  peek pp >>= fmap (DMatrix ctx) . newForeignPtr p_csr_free

-- | Krylov records: the state lives on the device.  The reference's steps are PURE (`iterate (bicgstabStep aa r0hat) st0 !! 20`,
--   README.md:208, keeps st0 alive), so `bicgstabStep` / `cgsStep` / `cgneStep` here clone the record (sla_krylov_clone) and advance the
--   clone; the `...InPlace` variants advance their argument for callers that own the record exclusively (no allocation per step).
newtype BICGSTAB = BICGSTAB (ForeignPtr SlaKrylov)
newtype CGS      = CGS (ForeignPtr SlaKrylov)

cloneK :: Ctx -> ForeignPtr SlaKrylov -> IO (ForeignPtr SlaKrylov)
cloneK ctx@(Ctx c) fs = withForeignPtr fs $ \ps -> alloca $ \pp -> do
  c_kry_clone c ps pp >>= check ctx "krylov_clone"
  peek pp >>= newForeignPtr p_kry_free

initWith :: String -> (Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> Ptr (Ptr SlaKrylov) -> IO Status)
         -> DMatrix -> DVector -> DVector -> IO (ForeignPtr SlaKrylov)
initWith who f (DMatrix ctx@(Ctx c) fa) (DVector _ fb) (DVector _ fx0) =
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> withForeignPtr fx0 $ \px -> alloca $ \pp -> do
    f c pa pb px pp >>= check ctx who
    peek pp >>= newForeignPtr p_kry_free

bicgsInit :: DMatrix -> DVector -> DVector -> IO BICGSTAB                       -- Sparse.hs:965-968
bicgsInit aa b x0 = BICGSTAB <$> initWith "bicgsInit" c_bicg_init aa b x0
cgsInit :: DMatrix -> DVector -> DVector -> IO CGS                              -- Sparse.hs:923-926
cgsInit aa b x0 = CGS <$> initWith "cgsInit" c_cgs_init aa b x0

bicgstabStepInPlace :: DMatrix -> DVector -> BICGSTAB -> IO BICGSTAB
bicgstabStepInPlace (DMatrix ctx@(Ctx c) fa) (DVector _ fr) st@(BICGSTAB fs) =
  withForeignPtr fa $ \pa -> withForeignPtr fr $ \pr -> withForeignPtr fs $ \ps -> c_bicg_step c pa pr ps >>= check ctx "bicgstabStep" >> return st
bicgstabStep :: DMatrix -> DVector -> BICGSTAB -> IO BICGSTAB                   -- Sparse.hs:970-981 (pure: the argument stays valid)
bicgstabStep aa@(DMatrix ctx _) r0hat (BICGSTAB fs) = cloneK ctx fs >>= bicgstabStepInPlace aa r0hat . BICGSTAB
cgsStepInPlace :: DMatrix -> DVector -> CGS -> IO CGS
cgsStepInPlace (DMatrix ctx@(Ctx c) fa) (DVector _ fr) st@(CGS fs) =
  withForeignPtr fa $ \pa -> withForeignPtr fr $ \pr -> withForeignPtr fs $ \ps -> c_cgs_step c pa pr ps >>= check ctx "cgsStep" >> return st
cgsStep :: DMatrix -> DVector -> CGS -> IO CGS                                  -- Sparse.hs:928-939 (pure)
cgsStep aa@(DMatrix ctx _) rhat (CGS fs) = cloneK ctx fs >>= cgsStepInPlace aa rhat . CGS

newtype CGNE = CGNE (ForeignPtr SlaKrylov)
cgneInit :: DMatrix -> DVector -> DVector -> IO CGNE                            -- Sparse.hs:862-866
cgneInit aa b x0 = CGNE <$> initWith "cgneInit" c_cgne_init aa b x0
cgneStepInPlace :: DMatrix -> CGNE -> IO CGNE                                   -- A^T is built once and cached on the device
cgneStepInPlace (DMatrix ctx@(Ctx c) fa) st@(CGNE fs) =
  withForeignPtr fa $ \pa -> withForeignPtr fs $ \ps -> c_cgne_step c pa ps >>= check ctx "cgneStep" >> return st
cgneStep :: DMatrix -> CGNE -> IO CGNE                                          -- Sparse.hs:868-878 (pure)
cgneStep aa@(DMatrix ctx _) (CGNE fs) = cloneK ctx fs >>= cgneStepInPlace aa . CGNE

-- | Preconditioners and triangular solves (Sparse.hs:673-721, 750-811).  The reference does not export the
--   preconditioners (Sparse.hs:17); the names are kept for the day it does.
wrapM :: Ctx -> Ptr (Ptr SlaCsr) -> IO DMatrix
wrapM ctx pp = peek pp >>= fmap (DMatrix ctx) . newForeignPtr p_csr_free

diagPartitions :: DMatrix -> IO (DMatrix, DMatrix, DMatrix)                    -- (sub-diagonal, diagonal, super-diagonal)
diagPartitions (DMatrix ctx@(Ctx c) fa) = withForeignPtr fa $ \pa -> alloca $ \pe -> alloca $ \pd -> alloca $ \pf -> do
  c_diag_parts c pa pe pd pf >>= check ctx "diagPartitions"
  (,,) <$> wrapM ctx pe <*> wrapM ctx pd <*> wrapM ctx pf

jacobiPre :: DMatrix -> IO DMatrix                                             -- recip <$> extractDiag x
jacobiPre (DMatrix ctx@(Ctx c) fa) = withForeignPtr fa $ \pa -> alloca $ \pm -> c_jacobi_pre c pa pm >>= check ctx "jacobiPre" >> wrapM ctx pm

mSsorPre :: DMatrix -> Double -> IO (DMatrix, DMatrix)                         -- (l, r), Sparse.hs:713-721
mSsorPre (DMatrix ctx@(Ctx c) fa) omega = withForeignPtr fa $ \pa -> alloca $ \pl -> alloca $ \pr -> do
  c_mssor_pre c pa omega pl pr >>= check ctx "mSsorPre"
  (,) <$> wrapM ctx pl <*> wrapM ctx pr

ilu0Pre :: DMatrix -> IO (DMatrix, DMatrix)                                   -- (l, u) with holes, Sparse.hs:696-706 (NeedsPivoting on a nearZero pivot)
ilu0Pre (DMatrix ctx@(Ctx c) fa) = withForeignPtr fa $ \pa -> alloca $ \pl -> alloca $ \pu -> do
  c_ilu0_pre c pa pl pu >>= check ctx "solveForLij"
  (,) <$> wrapM ctx pl <*> wrapM ctx pu

triSolveWith :: String -> (Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaVec -> Ptr SlaVec -> IO Status) -> DMatrix -> DVector -> IO DVector
triSolveWith who f (DMatrix ctx@(Ctx c) fa) b@(DVector _ fb) = do
  w@(DVector _ fw) <- dimD b >>= newVec ctx
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> withForeignPtr fw $ \pw -> f c pa pb pw >>= check ctx who
  return w

triLowerSolve, triUpperSolve :: DMatrix -> DVector -> IO DVector              -- Sparse.hs:750-778, 784-811 (NeedsPivoting on a nearZero diagonal)
triLowerSolve = triSolveWith "triLowerSolve" c_tri_lower
triUpperSolve = triSolveWith "triUpperSolve" c_tri_upper

data LinSolveMethod = GMRES_ | CGNE_ | BCG_ | CGS_ | BICGSTAB_ deriving (Eq, Show, Enum)   -- Sparse.hs:1007-1012

-- | linSolve0 method aa b x0 (Sparse.hs:1016-1072); NULL options = nits 200, tol = max 1e-6 (1e-4 * ||r0||), true residual.
linSolve0 :: LinSolveMethod -> DMatrix -> DVector -> DVector -> IO DVector
linSolve0 method (DMatrix ctx@(Ctx c) fa) b@(DVector _ fb) (DVector _ fx0) = do
  x@(DVector _ fx) <- dimD b >>= newVec ctx
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> withForeignPtr fx0 $ \p0 -> withForeignPtr fx $ \px ->
    alloca $ \pit -> alloca $ \pres ->
      c_linsolve0 c (fromIntegral (fromEnum method)) pa pb p0 nullPtr px pit pres >>= check ctx "linSolve0"
  return x

-- | aa <\> b : GMRES(30) from x0 = 0.1, as the reference's (commented-out) LinearSystem instance intended (Sparse.hs:1082-1088).
(<\>) :: DMatrix -> DVector -> IO DVector
(DMatrix ctx@(Ctx c) fa) <\> b@(DVector _ fb) = do
  n <- dimD b
  x0 <- toDeviceSV ctx (fromListDenseSV (fromIntegral n) (replicate (fromIntegral n) 0.1))
  x@(DVector _ fx) <- newVec ctx n
  let DVector _ f0 = x0
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> withForeignPtr f0 $ \p0 -> withForeignPtr fx $ \px ->
    alloca $ \pit -> alloca $ \pres -> c_gmres c pa pb p0 30 nullPtr px pit pres >>= check ctx "<\\>"
  return x

-- | arnoldi aa b kn (Sparse.hs:630-667): H is returned dense column-major ((nmax+1) x nmax); Q stays on the device behind a
--   ForeignPtr (released by sla_dense_free); `basisColumn` / `basisToHost` read it.
arnoldi :: DMatrix -> DVector -> Int -> IO (DDense, [Double], Int)
arnoldi (DMatrix ctx@(Ctx c) fa) (DVector _ fb) kn =
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> alloca $ \pq -> alloca $ \pn -> allocaArray ((kn + 1) * kn) $ \ph -> do
    st <- c_arnoldi c pa pb (fromIntegral kn) pq ph pn
    when (st /= 0 && st /= 5) (check ctx "arnoldi" st)      -- 5 = SLA_ERR_BREAKDOWN is informational
    nmax <- fromIntegral <$> peek pn
    h <- peekArray ((nmax + 1) * nmax) ph
    q <- peek pq >>= newForeignPtr p_dense_free
    return (DDense ctx q, h, nmax)

denseDims :: DDense -> IO (Int, Int)
denseDims (DDense _ fq) = withForeignPtr fq $ \pq -> alloca $ \pr -> alloca $ \pc -> do
  _ <- c_dense_dims pq pr pc
  (,) <$> (fromIntegral <$> peek pr) <*> (fromIntegral <$> peek pc)

-- | column j of the Arnoldi basis as a device vector (extractCol q j, SpMatrix.hs:329-337)
basisColumn :: DDense -> Int -> IO DVector
basisColumn (DDense ctx@(Ctx c) fq) j = withForeignPtr fq $ \pq -> alloca $ \pp -> do
  c_dense_col c pq (fromIntegral j) pp >>= check ctx "extractCol"
  peek pp >>= fmap (DVector ctx) . newForeignPtr p_vec_free

-- | the whole basis, column-major (n x (nmax + 1))
basisToHost :: DDense -> IO [Double]
basisToHost q@(DDense ctx@(Ctx c) fq) = do
  (r, cc) <- denseDims q
  withForeignPtr fq $ \pq -> allocaArray (r * cc) $ \po -> c_dense_to c pq po >>= check ctx "basisToHost" >> peekArray (r * cc) po

-- ---- the rest of MatrixRing (Class.hs:195-207; instance SpMatrix.hs:751-773) for a DENSE right operand -----------------------
-- dense operands are row-major blocks; elementType 0 = fp64 (bit-identical to matMat_), 1 = bf16 (BASELINE config 5)

toDeviceDense :: Ctx -> Int -> Int -> [Double] -> Int -> IO DDense                -- rows, cols, row-major entries, element type
toDeviceDense ctx@(Ctx c) r cc xs ty = withArray xs $ \px -> alloca $ \pp -> do
  c_dense_from c (fromIntegral r) (fromIntegral cc) px (fromIntegral ty) pp >>= check ctx "toDeviceDense"
  peek pp >>= fmap (DDense ctx) . newForeignPtr p_dense_free

fromDeviceDense :: DDense -> IO [Double]                                          -- row-major
fromDeviceDense d@(DDense ctx@(Ctx c) fd) = do
  (r, cc) <- denseDims d
  withForeignPtr fd $ \pd -> allocaArray (r * cc) $ \po -> c_dense_to_rm c pd po >>= check ctx "fromDeviceDense" >> peekArray (r * cc) po

mmWith :: String -> (Ptr SlaCtx -> Ptr SlaCsr -> Ptr SlaDense -> Ptr SlaDense -> IO Status) -> Int64 -> Int64 -> DMatrix -> DDense -> IO DDense
mmWith who f rows cols (DMatrix ctx@(Ctx c) fa) (DDense _ fb) = alloca $ \pp -> do
  c_dense_new c rows cols 0 pp >>= check ctx who
  fc <- peek pp >>= newForeignPtr p_dense_free
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> withForeignPtr fc $ \pc -> f c pa pb pc >>= check ctx who
  return (DDense ctx fc)

infixr 7 ##, ##^, #^#
(##), (##^), (#^#) :: DMatrix -> DDense -> IO DDense
aa ## b  = do { (m, _) <- dimM aa; (_, k) <- denseDims b; mmWith "matMat" c_spmm     m (fromIntegral k) aa b }    -- SpMatrix.hs:768-811
aa ##^ b = do { (m, _) <- dimM aa; (k, _) <- denseDims b; mmWith "matMat" c_spmm_abt m (fromIntegral k) aa b }    -- aa ## transpose b ; b is k x n
aa #^# b = do { (_, n) <- dimM aa; (_, k) <- denseDims b; mmWith "matMat" c_spmm_atb n (fromIntegral k) aa b }    -- transpose aa ## b (Class.hs:202)

normFrobenius :: DMatrix -> IO Double                                              -- sqrt (trace (m ##^ m)), SpMatrix.hs:751-752
normFrobenius (DMatrix ctx@(Ctx c) fa) = withForeignPtr fa $ \pa -> alloca $ \po -> c_frob c pa po >>= check ctx "normFrobenius" >> peek po

-- ---- one process, several GPUs (include/sla_b200.h: sla_init_multi) ------------------------------------------------------------
-- The library owns one worker thread and one context per GPU; matrices and vectors are GLOBAL (row-partitioned inside).
data SlaMCtx; data SlaMCsr; data SlaMVec
newtype MCtx  = MCtx (Ptr SlaMCtx)
data MMatrix  = MMatrix MCtx (ForeignPtr SlaMCsr)
data MVector  = MVector MCtx (ForeignPtr SlaMVec)

foreign import ccall safe "sla_init_multi"          c_minit     :: CInt -> Ptr CInt -> Ptr (Ptr SlaMCtx) -> IO Status
foreign import ccall safe "sla_finalize_multi"      c_mfinalize :: Ptr SlaMCtx -> IO ()
foreign import ccall safe "sla_multi_last_error"    c_mlast     :: Ptr SlaMCtx -> IO CString
foreign import ccall safe "sla_multi_csr_from_csr"  c_mfrom_csr :: Ptr SlaMCtx -> Int64 -> Int64 -> Int64 -> Ptr Int32 -> Ptr Int32 -> Ptr Double -> Ptr (Ptr SlaMCsr) -> IO Status
foreign import ccall safe "&sla_multi_csr_free"     p_mcsr_free :: FunPtr (Ptr SlaMCsr -> IO ())
foreign import ccall safe "sla_multi_vec_from_host" c_mvec_from :: Ptr SlaMCtx -> Int64 -> Ptr Double -> Ptr (Ptr SlaMVec) -> IO Status
foreign import ccall safe "sla_multi_vec_create"    c_mvec_new  :: Ptr SlaMCtx -> Int64 -> Ptr (Ptr SlaMVec) -> IO Status
foreign import ccall safe "sla_multi_vec_to_host"   c_mvec_to   :: Ptr SlaMCtx -> Ptr SlaMVec -> Ptr Double -> IO Status
foreign import ccall safe "sla_multi_vec_dim"       c_mvec_dim  :: Ptr SlaMVec -> IO Int64
foreign import ccall safe "&sla_multi_vec_free"     p_mvec_free :: FunPtr (Ptr SlaMVec -> IO ())
foreign import ccall safe "sla_multi_spmv"          c_mspmv     :: Ptr SlaMCtx -> Ptr SlaMCsr -> Ptr SlaMVec -> Ptr SlaMVec -> IO Status
foreign import ccall safe "sla_multi_dot"           c_mdot      :: Ptr SlaMCtx -> Ptr SlaMVec -> Ptr SlaMVec -> Ptr Double -> IO Status
foreign import ccall safe "sla_multi_linsolve0"     c_mlinsolve :: Ptr SlaMCtx -> CInt -> Ptr SlaMCsr -> Ptr SlaMVec -> Ptr SlaMVec -> Ptr () -> Ptr SlaMVec -> Ptr CInt -> Ptr Double -> IO Status
foreign import ccall safe "sla_multi_gmres"         c_mgmres    :: Ptr SlaMCtx -> Ptr SlaMCsr -> Ptr SlaMVec -> Ptr SlaMVec -> CInt -> Ptr () -> Ptr SlaMVec -> Ptr CInt -> Ptr Double -> IO Status

mcheck :: MCtx -> String -> Status -> IO ()
mcheck (MCtx m) who st = when (st /= 0) $ do
  msg <- c_mlast m >>= peekCString
  case st of
    1 -> throwIO (MatVecSizeMismatchException who (0, 0) 0)
    3 -> throwIO (IterE who msg :: IterationException ())
    _ -> ioError (userError (who ++ ": " ++ msg))

-- | withB200Multi n: GPUs 0 .. n-1 of this box, driven from this one process
withB200Multi :: Int -> (MCtx -> IO a) -> IO a
withB200Multi n = bracket open (\(MCtx m) -> c_mfinalize m)
  where open = alloca $ \pp -> do { st <- c_minit (fromIntegral n) nullPtr pp; m <- peek pp
                                  ; when (st /= 0) (ioError (userError "sla_init_multi failed")); return (MCtx m) }

-- | Marshal an SpMatrix as a global CSR (ascending (row, col) walk of immSM; every row present): the library cuts it into row blocks.
toDeviceSMMulti :: MCtx -> SpMatrix Double -> IO MMatrix
toDeviceSMMulti mc@(MCtx m) sm = do
  let nr = nrows sm
      rowsL = [ maybe [] I.toList (I.lookup i (immSM sm)) | i <- [0 .. nr - 1] ]
      rp = VS.fromList (scanl (+) 0 [ fromIntegral (length r) | r <- rowsL ]) :: VS.Vector Int32
      cs = VS.fromList [ fromIntegral j | r <- rowsL, (j, _) <- r ] :: VS.Vector Int32
      vs = VS.fromList [ x | r <- rowsL, (_, x) <- r ]
  VS.unsafeWith rp $ \prp -> VS.unsafeWith cs $ \pcs -> VS.unsafeWith vs $ \pvs -> alloca $ \pp -> do
    c_mfrom_csr m (fromIntegral nr) (fromIntegral (ncols sm)) (fromIntegral (VS.length vs)) prp pcs pvs pp >>= mcheck mc "fromListSM"
    peek pp >>= fmap (MMatrix mc) . newForeignPtr p_mcsr_free

toDeviceSVMulti :: MCtx -> SpVector Double -> IO MVector
toDeviceSVMulti mc@(MCtx m) v = VS.unsafeWith (VS.fromList (toDenseListSV v)) $ \px -> alloca $ \pp -> do
  c_mvec_from m (fromIntegral (dim v)) px pp >>= mcheck mc "toDeviceSV"
  peek pp >>= fmap (MVector mc) . newForeignPtr p_mvec_free

fromDeviceSVMulti :: MVector -> IO (SpVector Double)
fromDeviceSVMulti (MVector mc@(MCtx m) fv) = withForeignPtr fv $ \pv -> do
  n <- fromIntegral <$> c_mvec_dim pv
  allocaArray n $ \px -> c_mvec_to m pv px >>= mcheck mc "fromDeviceSV" >> (fromListDenseSV n <$> peekArray n px)

newMVec :: MCtx -> Int64 -> IO MVector
newMVec mc@(MCtx m) n = alloca $ \pp -> c_mvec_new m n pp >>= mcheck mc "zeroSV" >> peek pp >>= fmap (MVector mc) . newForeignPtr p_mvec_free

matVecMulti :: MMatrix -> MVector -> IO MVector                                     -- aa #> v across the GPUs
matVecMulti (MMatrix mc@(MCtx m) fa) (MVector _ fx) = do
  n <- withForeignPtr fx c_mvec_dim
  y@(MVector _ fy) <- newMVec mc n
  withForeignPtr fa $ \pa -> withForeignPtr fx $ \px -> withForeignPtr fy $ \py -> c_mspmv m pa px py >>= mcheck mc "matVec"
  return y

dotMulti :: MVector -> MVector -> IO Double
dotMulti (MVector mc@(MCtx m) fx) (MVector _ fy) =
  withForeignPtr fx $ \px -> withForeignPtr fy $ \py -> alloca $ \po -> c_mdot m px py po >>= mcheck mc "<.>" >> peek po

linSolve0Multi :: LinSolveMethod -> MMatrix -> MVector -> MVector -> IO MVector
linSolve0Multi method (MMatrix mc@(MCtx m) fa) (MVector _ fb) (MVector _ fx0) = do
  n <- withForeignPtr fb c_mvec_dim
  x@(MVector _ fx) <- newMVec mc n
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> withForeignPtr fx0 $ \p0 -> withForeignPtr fx $ \px ->
    alloca $ \pit -> alloca $ \pres -> c_mlinsolve m (fromIntegral (fromEnum method)) pa pb p0 nullPtr px pit pres >>= mcheck mc "linSolve0"
  return x

gmresMulti :: MMatrix -> MVector -> MVector -> Int -> IO MVector                     -- restarted GMRES(restart) across the GPUs
gmresMulti (MMatrix mc@(MCtx m) fa) (MVector _ fb) (MVector _ fx0) restart = do
  n <- withForeignPtr fb c_mvec_dim
  x@(MVector _ fx) <- newMVec mc n
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> withForeignPtr fx0 $ \p0 -> withForeignPtr fx $ \px ->
    alloca $ \pit -> alloca $ \pres -> c_mgmres m pa pb p0 (fromIntegral restart) nullPtr px pit pres >>= mcheck mc "gmres"
  return x
