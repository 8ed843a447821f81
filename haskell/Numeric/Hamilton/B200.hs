{-# LANGUAGE BangPatterns #-}
{-# LANGUAGE DataKinds #-}
{-# LANGUAGE DeriveGeneric #-}
{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE GADTs #-}
{-# LANGUAGE KindSignatures #-}
{-# LANGUAGE RankNTypes #-}
{-# LANGUAGE ScopedTypeVariables #-}
{-# LANGUAGE StandaloneDeriving #-}
{-# LANGUAGE TypeApplications #-}
{-# LANGUAGE TypeOperators #-}

-- |
-- Module      : Numeric.Hamilton.B200
-- Description : Drop-in replacement for the hot path of "Numeric.Hamilton" backed by
--               libhamilton_b200.so (include/hamilton_b200.h).
--
-- Same export list and the same types as @src\/Numeric\/Hamilton.hs:28-70@ of mstksg\/hamilton, plus batched and
-- multi-GPU additions.  Build: add @extra-libraries: hamilton_b200@ (and the library's directory to
-- @extra-lib-dirs@) to the cabal stanza; dependencies are the reference's own (hmatrix, vector-sized) minus @ad@ and
-- @hmatrix-gsl@.
--
-- THIS FILE HAS NEVER BEEN COMPILED: the image this repository is built in has no GHC (no ghc, cabal, stack or nix;
-- DESIGN.md section 5).  It is complete — every export is implemented, nothing is left @undefined@ — and
-- @tests\/test_cpu_host.py::test_haskell_shim_matches_header@ checks every @foreign import@ below against the C header
-- (name, arity, argument kinds), but a maintainer should expect the usual first-compile fixes.
--
-- How 'mkSystem' crosses a C boundary: its arguments are rank-2 polymorphic
-- (@forall a. RealFloat a => Vector n a -> Vector m a@, src\/Numeric\/Hamilton.hs:212-215), so instead of handing them
-- to the @ad@ package we instantiate @a@ at 'Tr', a number type whose arithmetic appends nodes to a Wengert list
-- (tape).  The tape goes to @hb_system_from_tape@, which differentiates it symbolically (what
-- @jacobianT@\/@hessianF@\/@grad@ do at run time, :221-224) and compiles specialised sm_100a code with NVRTC.
module Numeric.Hamilton.B200
  ( -- * Systems and states (the reference's export list)
    System, mkSystem, mkSystem', underlyingPos
  , Config (..), Phase (..), toPhase, fromPhase
  , momenta, velocities, keC, keP, pe, lagrangian, hamiltonian, hamEqs
  , stepHam, evolveHam, evolveHam'
  , stepHamC, evolveHamC, evolveHamC'
    -- * Batched additions (no counterpart in the reference)
  , Integrator (..), batchStep, batchEvolve, batchHamEqs
    -- * Multi-GPU ensembles (one process, all GPUs; hb_ensemble_*)
  , Ensemble, newEnsemble, ensembleInitRandom, ensembleUpload, ensembleStep, ensembleGather
  ) where

import Control.Concurrent.MVar
import Control.Exception (bracket_, evaluate)
import Control.Monad (forM, when)
import Data.IORef
import Data.Int
import Data.Kind (Type)
import Data.Proxy
import qualified Data.Vector.Sized as V
import qualified Data.Vector.Storable as VS
import Foreign
import Foreign.C.String
import GHC.Generics (Generic)
import GHC.TypeLits
import Numeric.LinearAlgebra.Static (R)
import qualified Numeric.LinearAlgebra.Static as H
import System.IO.Unsafe (unsafePerformIO)

-- ---------------------------------------------------------------------------------------------
-- C ABI (include/hamilton_b200.h).  Every foreign import names the reference definition it replaces.

data HbSystem
data HbEnsemble

-- | struct hb_op { int32 op, a, b, _pad; double c; }  (24 bytes)
data HbOp = HbOp !Int32 !Int32 !Int32 !Double

instance Storable HbOp where
  sizeOf _ = 24
  alignment _ = 8
  peek p = HbOp <$> peekByteOff p 0 <*> peekByteOff p 4 <*> peekByteOff p 8 <*> peekByteOff p 16
  poke p (HbOp o a b c) = do
    pokeByteOff p 0 o; pokeByteOff p 4 a; pokeByteOff p 8 b; pokeByteOff p 12 (0 :: Int32); pokeByteOff p 16 c

-- | struct hb_tape { int32 n_in; int32 n_ops; const hb_op* ops; int32 n_out; const int32* outs; }  (32 bytes)
withTape :: Int -> [HbOp] -> [Int32] -> (Ptr () -> IO a) -> IO a
withTape nIn ops outs k =
  withArrayLen ops $ \nOps pOps -> withArrayLen outs $ \nOut pOut -> allocaBytes 32 $ \t -> do
    pokeByteOff t 0 (fromIntegral nIn :: Int32); pokeByteOff t 4 (fromIntegral nOps :: Int32)
    pokeByteOff t 8 pOps; pokeByteOff t 16 (fromIntegral nOut :: Int32); pokeByteOff t 24 pOut
    k t

foreign import ccall safe "hb_last_error" c_last_error :: IO CString
-- mkSystem / mkSystem' (:201-254)
foreign import ccall safe "hb_system_from_tape"
  c_system_from_tape :: Int32 -> Int32 -> Ptr Double -> Ptr () -> Ptr () -> Int32 -> Ptr Double -> Int32 -> Ptr (Ptr HbSystem) -> IO Int32
foreign import ccall safe "&hb_system_free" p_system_free :: FunPtr (Ptr HbSystem -> IO ())
foreign import ccall safe "hb_underlying_pos" c_underlying_pos :: Ptr HbSystem -> Ptr Double -> Ptr Double -> IO Int32                   -- :174-178
foreign import ccall safe "hb_pe" c_pe :: Ptr HbSystem -> Ptr Double -> Ptr Double -> IO Int32                                            -- :182-186
foreign import ccall safe "hb_momenta" c_momenta :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32                   -- :262-269
foreign import ccall safe "hb_velocities" c_velocities :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32             -- :316-324
foreign import ccall safe "hb_ke_c" c_ke_c :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32                         -- :288-296
foreign import ccall safe "hb_ke_p" c_ke_p :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32                         -- :341-349
foreign import ccall safe "hb_lagrangian" c_lagrangian :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32             -- :301-309
foreign import ccall safe "hb_hamiltonian" c_hamiltonian :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32           -- :353-361
foreign import ccall safe "hb_ham_eqs" c_ham_eqs :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32     -- :370-387
foreign import ccall safe "hb_step_ham" c_step_ham :: Ptr HbSystem -> Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32     -- :390-402
foreign import ccall safe "hb_evolve_ham" c_evolve_ham :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> Int32 -> Ptr Double -> IO Int32   -- :433-462
foreign import ccall safe "hb_step_ham_c" c_step_ham_c :: Ptr HbSystem -> Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32 -- :505-515
foreign import ccall safe "hb_evolve_ham_c" c_evolve_ham_c :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> Int32 -> Ptr Double -> IO Int32 -- :488-498
-- batched (additive)
foreign import ccall safe "hb_batch_ham_eqs"
  c_batch_ham_eqs :: Ptr HbSystem -> Int64 -> Int32 -> Int32 -> Ptr Double -> Ptr Double -> Ptr Int32 -> Ptr () -> IO Int32
foreign import ccall safe "hb_batch_step"
  c_batch_step :: Ptr HbSystem -> Int32 -> Double -> Int32 -> Int64 -> Int32 -> Int32 -> Ptr Double -> Ptr Double -> Ptr Int32 -> Ptr () -> IO Int32
foreign import ccall safe "hb_batch_evolve"
  c_batch_evolve :: Ptr HbSystem -> Int32 -> Int32 -> Int64 -> Int32 -> Int32 -> Ptr Double -> Ptr Double -> Int32 -> Ptr Double -> Ptr Int32 -> Ptr () -> IO Int32
-- multi-GPU ensembles (additive)
foreign import ccall safe "hb_ensemble_create"
  c_ensemble_create :: Ptr HbSystem -> Int32 -> Ptr Int32 -> Int64 -> Ptr (Ptr HbEnsemble) -> IO Int32
foreign import ccall safe "&hb_ensemble_free" p_ensemble_free :: FunPtr (Ptr HbEnsemble -> IO ())
foreign import ccall safe "hb_ensemble_init_random" c_ensemble_init_random :: Ptr HbEnsemble -> Word64 -> Ptr Double -> Ptr Double -> IO Int32
foreign import ccall safe "hb_ensemble_upload" c_ensemble_upload :: Ptr HbEnsemble -> Ptr Double -> IO Int32
foreign import ccall safe "hb_ensemble_step" c_ensemble_step :: Ptr HbEnsemble -> Int32 -> Double -> Int32 -> Int32 -> Ptr Double -> IO Int32
foreign import ccall safe "hb_ensemble_gather" c_ensemble_gather :: Ptr HbEnsemble -> Ptr Double -> Ptr Double -> IO Int32

-- | Non-zero status -> 'error', mirroring the reference's partiality (:425, :462, hmatrix's @inv@ exception).
orDie :: String -> IO Int32 -> IO ()
orDie what act = do
  rc <- act
  when (rc /= 0) $ c_last_error >>= peekCString >>= \m -> error ("Numeric.Hamilton.B200." ++ what ++ ": " ++ m)

-- ---------------------------------------------------------------------------------------------
-- The tracing number type (the stand-in for ad's dual / tower numbers).
--
-- A 'Tr' is either a literal (no tape needed, so 'fromInteger', 'fromRational' and 'pi' are total) or the index of a
-- node on the tape being recorded.  Exactly one trace is recorded at a time ('traceLock'); all nodes of a trace are
-- forced inside 'traceFn', before the tape is closed.  Literal-only arithmetic is folded in Haskell.

data Tr = Lit !Double | Node !Int32

-- | The tape under construction (reversed) and its length.
data Tape = Tape ![HbOp] !Int32

currentTape :: IORef Tape
currentTape = unsafePerformIO (newIORef (Tape [] 0))
{-# NOINLINE currentTape #-}

traceLock :: MVar ()
traceLock = unsafePerformIO (newMVar ())
{-# NOINLINE traceLock #-}

-- | Appends one node; the operands were forced by the callers ('nodeOf' is strict), so they sit earlier on the tape
-- ("node k may only refer to nodes < k", include/hamilton_b200.h).
push :: HbOp -> Int32
push !o = unsafePerformIO $ atomicModifyIORef' currentTape $ \(Tape os k) -> (Tape (o : os) (k + 1), k)
{-# NOINLINE push #-}

-- | Index of a value on the tape; literals are materialised as HB_OP_CONST nodes.
nodeOf :: Tr -> Int32
nodeOf (Node i) = i
nodeOf (Lit c) = push (HbOp 1 0 0 c)
{-# NOINLINE nodeOf #-}

-- hb_opcode numbering (include/hamilton_b200.h): INPUT=0 CONST=1 PARAM=2 ADD=3 SUB=4 MUL=5 DIV=6 NEG=7 RECIP=8 ABS=9
-- SIGNUM=10 SQRT=11 EXP=12 LOG=13 SIN=14 COS=15 TAN=16 ASIN=17 ACOS=18 ATAN=19 SINH=20 COSH=21 TANH=22 ASINH=23
-- ACOSH=24 ATANH=25 POW=26 POWI=27 ATAN2=28
bin :: Int32 -> (Double -> Double -> Double) -> Tr -> Tr -> Tr
bin _ f (Lit a) (Lit b) = Lit (f a b)
bin op _ a b = let !ia = nodeOf a; !ib = nodeOf b in Node (push (HbOp op ia ib 0))
{-# NOINLINE bin #-}

un :: Int32 -> (Double -> Double) -> Tr -> Tr
un _ f (Lit a) = Lit (f a)
un op _ a = let !ia = nodeOf a in Node (push (HbOp op ia 0 0))
{-# NOINLINE un #-}
-- (NOINLINE: a 'push' must never be specialised to constant operands and floated out of the function being traced —
-- it would be evaluated once and its node index reused by a later trace.  Every traced node depends on an input of the
-- traced function, and the function is re-evaluated at every 'traceFn'.)

instance Num Tr where
  (+) = bin 3 (+); (-) = bin 4 (-); (*) = bin 5 (*)
  negate = un 7 negate; abs = un 9 abs; signum = un 10 signum
  fromInteger = Lit . fromInteger
instance Fractional Tr where
  (/) = bin 6 (/); recip = un 8 recip; fromRational = Lit . fromRational
instance Floating Tr where
  pi = Lit pi
  exp = un 12 exp; log = un 13 log; sqrt = un 11 sqrt; sin = un 14 sin; cos = un 15 cos; tan = un 16 tan
  asin = un 17 asin; acos = un 18 acos; atan = un 19 atan; sinh = un 20 sinh; cosh = un 21 cosh; tanh = un 22 tanh
  asinh = un 23 asinh; acosh = un 24 acosh; atanh = un 25 atanh
  (**) = bin 26 (**)
  logBase b x = log x / log b

-- | A tape records straight-line arithmetic.  Anything that needs the VALUE of a traced number (comparisons, rounding,
-- decoding) would be data-dependent control flow, which neither this tape nor symbolic differentiation can represent;
-- on literals these operations work, on traced values they raise this error.  None of the reference's example systems
-- needs them (app/Examples.hs:61-183).
untraceable :: String -> a
untraceable what = error ("Numeric.Hamilton.B200: `" ++ what ++ "` on a traced coordinate is data-dependent control flow; "
                          ++ "mkSystem's functions must be straight-line arithmetic (literals are fine)")

instance Eq Tr where
  Lit a == Lit b = a == b
  _ == _ = untraceable "=="
instance Ord Tr where
  compare (Lit a) (Lit b) = compare a b
  compare _ _ = untraceable "compare"
instance Real Tr where
  toRational (Lit a) = toRational a
  toRational _ = untraceable "toRational"
instance RealFrac Tr where
  properFraction (Lit a) = let (k, f) = properFraction a in (k, Lit f)
  properFraction _ = untraceable "properFraction"
instance RealFloat Tr where
  floatRadix _ = floatRadix (0 :: Double)
  floatDigits _ = floatDigits (0 :: Double)
  floatRange _ = floatRange (0 :: Double)
  isIEEE _ = True
  decodeFloat (Lit a) = decodeFloat a
  decodeFloat _ = untraceable "decodeFloat"
  encodeFloat m e = Lit (encodeFloat m e)
  isNaN (Lit a) = isNaN a
  isNaN _ = untraceable "isNaN"
  isInfinite (Lit a) = isInfinite a
  isInfinite _ = untraceable "isInfinite"
  isDenormalized (Lit a) = isDenormalized a
  isDenormalized _ = untraceable "isDenormalized"
  isNegativeZero (Lit a) = isNegativeZero a
  isNegativeZero _ = untraceable "isNegativeZero"
  atan2 = bin 28 atan2

-- | Records @f@ applied to @nIn@ fresh inputs; returns the tape and the node index of every output.
traceFn :: Int -> ([Tr] -> [Tr]) -> IO ([HbOp], [Int32])
traceFn nIn f = bracket_ (takeMVar traceLock) (putMVar traceLock ()) $ do
  writeIORef currentTape (Tape [] 0)
  ins <- forM [0 .. nIn - 1] $ \j -> evaluate (Node (push (HbOp 0 (fromIntegral j) 0 0)))
  outs <- forM (f ins) $ \o -> evaluate (nodeOf o)          -- forces every node of the trace, in dependency order
  Tape ops _ <- readIORef currentTape
  return (reverse ops, outs)

-- ---------------------------------------------------------------------------------------------
-- Public API: same types as the reference

-- | src/Numeric/Hamilton.hs:103-113
data Config :: Nat -> Type where
  Cfg :: { cfgPositions :: !(R n), cfgVelocities :: !(R n) } -> Config n
  deriving (Generic)
deriving instance KnownNat n => Show (Config n)

-- | src/Numeric/Hamilton.hs:133-143
data Phase :: Nat -> Type where
  Phs :: { phsPositions :: !(R n), phsMomenta :: !(R n) } -> Phase n
  deriving (Generic)
deriving instance KnownNat n => Show (Phase n)

-- | src/Numeric/Hamilton.hs:160-169: abstract in the reference too (constructor not exported), so the representation
-- is free: a handle to the compiled system inside the library, released by the garbage collector.
newtype System (m :: Nat) (n :: Nat) = Sys (ForeignPtr HbSystem)

data Integrator = RK4 | RKF45_GSL deriving (Eq, Show, Enum)

natInt :: forall k. KnownNat k => Proxy k -> Int
natInt = fromIntegral . natVal

mkSystemWith :: forall m n. (KnownNat m, KnownNat n) => Bool -> R m -> ([Tr] -> [Tr]) -> ([Tr] -> Tr) -> System m n
mkSystemWith onCart inertia f u = unsafePerformIO $ do
  let m = natInt (Proxy @m); n = natInt (Proxy @n); nu = if onCart then m else n
  (fops, fouts) <- traceFn n f
  (uops, uouts) <- traceFn nu (\xs -> [u xs])
  withTape n fops fouts $ \ft -> withTape nu uops uouts $ \ut ->
    VS.unsafeWith (H.extract inertia) $ \pw -> alloca $ \out -> do
      orDie "mkSystem" $ c_system_from_tape (fromIntegral m) (fromIntegral n) pw ft ut (if onCart then 1 else 0) nullPtr 0 out
      Sys <$> (peek out >>= newForeignPtr p_system_free)
{-# NOINLINE mkSystemWith #-}

-- | The rank-2 arguments instantiated at the tracing type, on plain lists.
atTr :: forall j k. (KnownNat j, KnownNat k) => (forall a. RealFloat a => V.Vector j a -> V.Vector k a) -> [Tr] -> [Tr]
atTr f xs = case V.fromList xs :: Maybe (V.Vector j Tr) of
  Just v -> V.toList (f v)
  Nothing -> error "Numeric.Hamilton.B200: internal error (input arity)"

atTr1 :: forall j. KnownNat j => (forall a. RealFloat a => V.Vector j a -> a) -> [Tr] -> Tr
atTr1 u xs = case V.fromList xs :: Maybe (V.Vector j Tr) of
  Just v -> u v
  Nothing -> error "Numeric.Hamilton.B200: internal error (input arity)"

-- | src/Numeric/Hamilton.hs:201-225: potential on the generalized coordinates.
mkSystem
  :: forall m n. (KnownNat m, KnownNat n)
  => R m
  -> (forall a. RealFloat a => V.Vector n a -> V.Vector m a)
  -> (forall a. RealFloat a => V.Vector n a -> a)
  -> System m n
mkSystem w f u = mkSystemWith False w (atTr @n @m f) (atTr1 @n u)

-- | src/Numeric/Hamilton.hs:238-254: potential on the underlying Cartesian coordinates (the library forms @u . f@).
mkSystem'
  :: forall m n. (KnownNat m, KnownNat n)
  => R m
  -> (forall a. RealFloat a => V.Vector n a -> V.Vector m a)
  -> (forall a. RealFloat a => V.Vector m a -> a)
  -> System m n
mkSystem' w f u = mkSystemWith True w (atTr @n @m f) (atTr1 @m u)

-- ---- marshalling helpers ---------------------------------------------------------------------

withR :: KnownNat k => R k -> (Ptr Double -> IO a) -> IO a
withR v = VS.unsafeWith (H.extract v)       -- extract expands hmatrix's compact constant vectors

readR :: forall k. KnownNat k => Ptr Double -> IO (R k)
readR p = H.vector <$> peekArray (natInt (Proxy @k)) p

-- | @f sys a b out@ with two @R@ inputs and one @R@ output.
call21 :: forall m n i j o. (KnownNat i, KnownNat j, KnownNat o)
       => String -> (Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32) -> System m n -> R i -> R j -> R o
call21 what f (Sys fp) a b = unsafePerformIO $ withForeignPtr fp $ \s -> withR a $ \pa -> withR b $ \pb ->
  allocaArray (natInt (Proxy @o)) $ \po -> orDie what (f s pa pb po) >> readR po

-- | @f sys a b out@ with two @R@ inputs and one scalar output.
call2s :: (KnownNat i, KnownNat j)
       => String -> (Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32) -> System m n -> R i -> R j -> Double
call2s what f (Sys fp) a b = unsafePerformIO $ withForeignPtr fp $ \s -> withR a $ \pa -> withR b $ \pb ->
  alloca $ \po -> orDie what (f s pa pb po) >> peek po

-- ---- the reference's functions ---------------------------------------------------------------

-- | src/Numeric/Hamilton.hs:174-178
underlyingPos :: forall m n. (KnownNat m, KnownNat n) => System m n -> R n -> R m
underlyingPos (Sys fp) q = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq ->
  allocaArray (natInt (Proxy @m)) $ \px -> orDie "underlyingPos" (c_underlying_pos s pq px) >> readR px

-- | src/Numeric/Hamilton.hs:182-186
pe :: forall m n. (KnownNat m, KnownNat n) => System m n -> R n -> Double
pe (Sys fp) q = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq ->
  alloca $ \pu -> orDie "pe" (c_pe s pq pu) >> peek pu

-- | src/Numeric/Hamilton.hs:262-269
momenta :: (KnownNat m, KnownNat n) => System m n -> Config n -> R n
momenta s (Cfg q v) = call21 "momenta" c_momenta s q v

-- | src/Numeric/Hamilton.hs:279-284
toPhase :: (KnownNat m, KnownNat n) => System m n -> Config n -> Phase n
toPhase s c = Phs (cfgPositions c) (momenta s c)

-- | src/Numeric/Hamilton.hs:288-296
keC :: (KnownNat m, KnownNat n) => System m n -> Config n -> Double
keC s (Cfg q v) = call2s "keC" c_ke_c s q v

-- | src/Numeric/Hamilton.hs:301-309
lagrangian :: (KnownNat m, KnownNat n) => System m n -> Config n -> Double
lagrangian s (Cfg q v) = call2s "lagrangian" c_lagrangian s q v

-- | src/Numeric/Hamilton.hs:316-324 (an SPD solve in the library where the reference forms @inv jmj@)
velocities :: (KnownNat m, KnownNat n) => System m n -> Phase n -> R n
velocities s (Phs q p) = call21 "velocities" c_velocities s q p

-- | src/Numeric/Hamilton.hs:332-337
fromPhase :: (KnownNat m, KnownNat n) => System m n -> Phase n -> Config n
fromPhase s p = Cfg (phsPositions p) (velocities s p)

-- | src/Numeric/Hamilton.hs:341-349
keP :: (KnownNat m, KnownNat n) => System m n -> Phase n -> Double
keP s (Phs q p) = call2s "keP" c_ke_p s q p

-- | src/Numeric/Hamilton.hs:353-361
hamiltonian :: (KnownNat m, KnownNat n) => System m n -> Phase n -> Double
hamiltonian s (Phs q p) = call2s "hamiltonian" c_hamiltonian s q p

-- | src/Numeric/Hamilton.hs:370-387
hamEqs :: forall m n. (KnownNat m, KnownNat n) => System m n -> Phase n -> (R n, R n)
hamEqs (Sys fp) (Phs q p) = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq -> withR p $ \pp ->
  allocaArray n $ \pdq -> allocaArray n $ \pdp -> do
    orDie "hamEqs" (c_ham_eqs s pq pp pdq pdp)
    (,) <$> readR pdq <*> readR pdp
  where n = natInt (Proxy @n)

-- | src/Numeric/Hamilton.hs:390-402: the library runs the same GSL-semantics adaptive RKF45 solve over (0, r).
stepHam :: forall m n. (KnownNat m, KnownNat n) => Double -> System m n -> Phase n -> Phase n
stepHam r (Sys fp) (Phs q p) = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq -> withR p $ \pp ->
  allocaArray n $ \pqo -> allocaArray n $ \ppo -> do
    orDie "stepHam" (c_step_ham s r pq pp pqo ppo)
    Phs <$> readR pqo <*> readR ppo
  where n = natInt (Proxy @n)

-- | Rows of @s x 2n@ doubles -> @s@ values built from (first half, second half).
rowsOf :: forall n a. KnownNat n => (R n -> R n -> a) -> Int -> Ptr Double -> IO [a]
rowsOf mk s pout = forM [0 .. s - 1] $ \k -> do
  let n = natInt (Proxy @n); row = pout `advancePtr` (k * 2 * n)
  mk <$> readR row <*> readR (row `advancePtr` n)

evolveList :: forall m n. (KnownNat m, KnownNat n) => System m n -> Phase n -> [Double] -> [Phase n]
evolveList (Sys fp) (Phs q p) ts = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq -> withR p $ \pp ->
  withArrayLen ts $ \ns pts -> allocaArray (ns * 2 * natInt (Proxy @n)) $ \pout -> do
    orDie "evolveHam" (c_evolve_ham s pq pp pts (fromIntegral ns) pout)
    rowsOf @n Phs ns pout

-- | src/Numeric/Hamilton.hs:433-462: row 0 is the initial state; h and the FSAL derivative carry across grid points.
evolveHam :: forall m n s. (KnownNat m, KnownNat n, KnownNat s, 2 <= s) => System m n -> Phase n -> V.Vector s Double -> V.Vector s (Phase n)
evolveHam sys p0 ts = case V.fromList (evolveList sys p0 (V.toList ts)) of
  Just v -> v
  Nothing -> error "Numeric.Hamilton.B200.evolveHam: internal error (row count)"

-- | src/Numeric/Hamilton.hs:409-429: @[]@ gives @[]@; a single time @[x]@ means the grid @[0, x]@ without its first row.
evolveHam' :: forall m n. (KnownNat m, KnownNat n) => System m n -> Phase n -> [Double] -> [Phase n]
evolveHam' _ _ [] = []
evolveHam' sys p0 [x] = drop 1 (evolveList sys p0 [0, x])
evolveHam' sys p0 ts = evolveList sys p0 ts

-- | src/Numeric/Hamilton.hs:505-515
stepHamC :: forall m n. (KnownNat m, KnownNat n) => Double -> System m n -> Config n -> Config n
stepHamC r (Sys fp) (Cfg q v) = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq -> withR v $ \pv ->
  allocaArray n $ \pqo -> allocaArray n $ \pvo -> do
    orDie "stepHamC" (c_step_ham_c s r pq pv pqo pvo)
    Cfg <$> readR pqo <*> readR pvo
  where n = natInt (Proxy @n)

evolveListC :: forall m n. (KnownNat m, KnownNat n) => System m n -> Config n -> [Double] -> [Config n]
evolveListC (Sys fp) (Cfg q v) ts = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq -> withR v $ \pv ->
  withArrayLen ts $ \ns pts -> allocaArray (ns * 2 * natInt (Proxy @n)) $ \pout -> do
    orDie "evolveHamC" (c_evolve_ham_c s pq pv pts (fromIntegral ns) pout)
    rowsOf @n Cfg ns pout

-- | src/Numeric/Hamilton.hs:488-498
evolveHamC :: forall m n s. (KnownNat m, KnownNat n, KnownNat s, 2 <= s) => System m n -> Config n -> V.Vector s Double -> V.Vector s (Config n)
evolveHamC sys c0 ts = case V.fromList (evolveListC sys c0 (V.toList ts)) of
  Just v -> v
  Nothing -> error "Numeric.Hamilton.B200.evolveHamC: internal error (row count)"

-- | src/Numeric/Hamilton.hs:470-480 (same edge cases as 'evolveHam'')
evolveHamC' :: forall m n. (KnownNat m, KnownNat n) => System m n -> Config n -> [Double] -> [Config n]
evolveHamC' _ _ [] = []
evolveHamC' sys c0 [x] = drop 1 (evolveListC sys c0 [0, x])
evolveHamC' sys c0 ts = evolveListC sys c0 ts

-- ---------------------------------------------------------------------------------------------
-- Batched additions: N trajectories at once, a storable vector of N Phases laid out [q, p] per trajectory
-- (HB_LAYOUT_AOS = 0, HB_MEM_HOST = 0).  Numerical failures of single trajectories do not abort a batch; pass a flags
-- array through the C ABI directly if they must be inspected.

phaseWidth :: forall m n. KnownNat n => System m n -> Int
phaseWidth _ = 2 * natInt (Proxy @n)

-- | 'hamEqs' on every Phase of the batch.
batchHamEqs :: forall m n. (KnownNat m, KnownNat n) => System m n -> VS.Vector Double -> VS.Vector Double
batchHamEqs sys@(Sys fp) ys = unsafePerformIO $ withForeignPtr fp $ \s -> do
  let len = VS.length ys; nTraj = len `div` phaseWidth sys
  out <- mallocForeignPtrArray len
  VS.unsafeWith ys $ \pin -> withForeignPtr out $ \pout ->
    orDie "batchHamEqs" $ c_batch_ham_eqs s (fromIntegral nTraj) 0 0 pin pout nullPtr nullPtr
  return (VS.unsafeFromForeignPtr0 out len)

-- | @nsteps@ steps of size @dt@ on every Phase ('RK4': classical fixed step; 'RKF45_GSL': @nsteps@ x 'stepHam' @dt@).
batchStep :: forall m n. (KnownNat m, KnownNat n) => Integrator -> Double -> Int -> System m n -> VS.Vector Double -> VS.Vector Double
batchStep integ dt nsteps sys@(Sys fp) ys = unsafePerformIO $ withForeignPtr fp $ \s -> do
  let len = VS.length ys; nTraj = len `div` phaseWidth sys
  out <- mallocForeignPtrArray len
  VS.unsafeWith ys $ \pin -> withForeignPtr out $ \pout ->
    orDie "batchStep" $ c_batch_step s (fromIntegral (fromEnum integ)) dt (fromIntegral nsteps) (fromIntegral nTraj) 0 0 pin pout nullPtr nullPtr
  return (VS.unsafeFromForeignPtr0 out len)

-- | 'evolveHam' for every Phase of the batch over one shared time grid: the result holds one batch per grid time
-- (the first is the input).  @rk4Substeps@ equal RK4 steps per grid interval when the integrator is 'RK4'.
batchEvolve :: forall m n. (KnownNat m, KnownNat n) => Integrator -> Int -> System m n -> VS.Vector Double -> [Double] -> [VS.Vector Double]
batchEvolve integ rk4Substeps sys@(Sys fp) y0 ts = unsafePerformIO $ withForeignPtr fp $ \s -> do
  let len = VS.length y0; nTraj = len `div` phaseWidth sys; ns = length ts
  out <- mallocForeignPtrArray (len * ns)
  VS.unsafeWith y0 $ \pin -> withArray ts $ \pts -> withForeignPtr out $ \pout ->
    orDie "batchEvolve" $ c_batch_evolve s (fromIntegral (fromEnum integ)) (fromIntegral rk4Substeps) (fromIntegral nTraj) 0 0 pin pts (fromIntegral ns) pout nullPtr nullPtr
  let whole = VS.unsafeFromForeignPtr0 out (len * ns)
  return [VS.slice (k * len) len whole | k <- [0 .. ns - 1]]

-- ---------------------------------------------------------------------------------------------
-- Multi-GPU ensembles: N independent initial conditions block-split over the GPUs of this process (stepHam is a pure
-- function of each Phase, :390-399), one NCCL all-gather to collect them.  IO, because an ensemble is mutable state on
-- the devices.

data Ensemble (m :: Nat) (n :: Nat) = Ens (ForeignPtr HbEnsemble) (System m n) Int

-- | @newEnsemble sys ndev nTraj@ over devices 0 .. ndev-1.
newEnsemble :: System m n -> Int -> Int -> IO (Ensemble m n)
newEnsemble sys@(Sys fp) ndev nTraj = withForeignPtr fp $ \s -> alloca $ \out -> do
  orDie "newEnsemble" $ c_ensemble_create s (fromIntegral ndev) nullPtr (fromIntegral nTraj) out
  e <- peek out >>= newForeignPtr p_ensemble_free
  return (Ens e sys nTraj)

-- | Counter-based initial Phases, uniform in the box [lo, hi] (2n numbers each), generated on the devices.
ensembleInitRandom :: Ensemble m n -> Word64 -> [Double] -> [Double] -> IO ()
ensembleInitRandom (Ens e _ _) seed lo hi = withForeignPtr e $ \pe' -> withArray lo $ \plo -> withArray hi $ \phi ->
  orDie "ensembleInitRandom" (c_ensemble_init_random pe' seed plo phi)

-- | Initial Phases from the host: a storable vector of N Phases.
ensembleUpload :: Ensemble m n -> VS.Vector Double -> IO ()
ensembleUpload (Ens e _ _) ys = withForeignPtr e $ \pe' -> VS.unsafeWith ys $ \py -> orDie "ensembleUpload" (c_ensemble_upload pe' py)

-- | @launches@ x (@nsteps@ steps of size @dt@) on every shard concurrently; returns the device time in milliseconds.
ensembleStep :: Ensemble m n -> Integrator -> Double -> Int -> Int -> IO Double
ensembleStep (Ens e _ _) integ dt nsteps launches = withForeignPtr e $ \pe' -> alloca $ \pms -> do
  orDie "ensembleStep" (c_ensemble_step pe' (fromIntegral (fromEnum integ)) dt (fromIntegral nsteps) (fromIntegral launches) pms)
  peek pms

-- | One all-gather over NVLink; the N Phases come back as one storable vector.
ensembleGather :: forall m n. KnownNat n => Ensemble m n -> IO (VS.Vector Double)
ensembleGather (Ens e sys nTraj) = withForeignPtr e $ \pe' -> do
  let len = nTraj * phaseWidth sys
  out <- mallocForeignPtrArray len
  withForeignPtr out $ \pout -> orDie "ensembleGather" (c_ensemble_gather pe' pout nullPtr)
  return (VS.unsafeFromForeignPtr0 out len)
