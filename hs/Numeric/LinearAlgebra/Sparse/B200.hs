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
  ( Ctx, DMatrix, DVector, BICGSTAB(..), CGS(..), CGNE(..)
  , withB200, toDeviceSM, toDeviceSV, fromDeviceSV
  , (#>), (<#), (<.>), (^+^), (^-^), (.*), (./), norm2, normalize2, transpose
  , bicgsInit, bicgstabStep, cgsInit, cgsStep, cgneInit, cgneStep
  , LinSolveMethod(..), linSolve0, arnoldi, (<\>)
  , diagPartitions, jacobiPre, mSsorPre, triLowerSolve, triUpperSolve
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

-- | Krylov records: the state lives on the device and is advanced IN PLACE by the step functions.
newtype BICGSTAB = BICGSTAB (ForeignPtr SlaKrylov)
newtype CGS      = CGS (ForeignPtr SlaKrylov)

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

bicgstabStep :: DMatrix -> DVector -> BICGSTAB -> IO BICGSTAB                   -- Sparse.hs:970-981
bicgstabStep (DMatrix ctx@(Ctx c) fa) (DVector _ fr) st@(BICGSTAB fs) =
  withForeignPtr fa $ \pa -> withForeignPtr fr $ \pr -> withForeignPtr fs $ \ps -> c_bicg_step c pa pr ps >>= check ctx "bicgstabStep" >> return st
cgsStep :: DMatrix -> DVector -> CGS -> IO CGS                                  -- Sparse.hs:928-939
cgsStep (DMatrix ctx@(Ctx c) fa) (DVector _ fr) st@(CGS fs) =
  withForeignPtr fa $ \pa -> withForeignPtr fr $ \pr -> withForeignPtr fs $ \ps -> c_cgs_step c pa pr ps >>= check ctx "cgsStep" >> return st

newtype CGNE = CGNE (ForeignPtr SlaKrylov)
cgneInit :: DMatrix -> DVector -> DVector -> IO CGNE                            -- Sparse.hs:862-866
cgneInit aa b x0 = CGNE <$> initWith "cgneInit" c_cgne_init aa b x0
cgneStep :: DMatrix -> CGNE -> IO CGNE                                          -- Sparse.hs:868-878 (A^T is cached on the device)
cgneStep (DMatrix ctx@(Ctx c) fa) st@(CGNE fs) =
  withForeignPtr fa $ \pa -> withForeignPtr fs $ \ps -> c_cgne_step c pa ps >>= check ctx "cgneStep" >> return st

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

-- | arnoldi aa b kn (Sparse.hs:630-667): H is returned dense column-major ((nmax+1) x nmax); Q stays on the device.
arnoldi :: DMatrix -> DVector -> Int -> IO (Ptr SlaDense, [Double], Int)
arnoldi (DMatrix ctx@(Ctx c) fa) (DVector _ fb) kn =
  withForeignPtr fa $ \pa -> withForeignPtr fb $ \pb -> alloca $ \pq -> alloca $ \pn -> allocaArray ((kn + 1) * kn) $ \ph -> do
    st <- c_arnoldi c pa pb (fromIntegral kn) pq ph pn
    when (st /= 0 && st /= 5) (check ctx "arnoldi" st)      -- 5 = SLA_ERR_BREAKDOWN is informational
    nmax <- fromIntegral <$> peek pn
    h <- peekArray ((nmax + 1) * nmax) ph
    q <- peek pq
    return (q, h, nmax)
