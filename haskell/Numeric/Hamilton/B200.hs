{-# LANGUAGE DataKinds #-}
{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE KindSignatures #-}
{-# LANGUAGE RankNTypes #-}
{-# LANGUAGE ScopedTypeVariables #-}
{-# LANGUAGE TypeApplications #-}

-- |
-- Module      : Numeric.Hamilton.B200
-- Description : Drop-in replacement for the hot path of "Numeric.Hamilton" backed by
--               libhamilton_b200.so (include/hamilton_b200.h).
--
-- NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no GHC.  This is the binding a
-- maintainer of mstksg/hamilton adds (INTEGRATION.md walks through it); every foreign import below
-- names the reference definition it replaces (src/Numeric/Hamilton.hs line numbers).
--
-- The trick that makes 'mkSystem' work across a C boundary: its argument is rank-2 polymorphic
-- (@forall a. RealFloat a => Vector n a -> Vector m a@, src/Numeric/Hamilton.hs:212), so instead
-- of handing it to the @ad@ package we instantiate @a@ at 'Tr', a number type whose arithmetic
-- appends nodes to a Wengert list.  The list is shipped to 'hb_system_from_tape', which
-- differentiates it symbolically (what jacobianT/hessianF/grad do at run time, :221-224) and
-- compiles specialised sm_100a code with NVRTC.
module Numeric.Hamilton.B200
  ( System, mkSystem, mkSystem'
  , Config (..), Phase (..)
  , underlyingPos, pe, momenta, toPhase, keC, lagrangian, velocities, fromPhase, keP, hamiltonian
  , hamEqs, stepHam, evolveHam, evolveHam', stepHamC, evolveHamC, evolveHamC'
    -- * Batched additions (no counterpart in the reference)
  , batchStep, Integrator (..)
  ) where

import Control.Monad (forM_, when)
import Data.IORef
import Data.Proxy
import qualified Data.Vector.Sized as V
import qualified Data.Vector.Storable as VS
import Foreign
import Foreign.C.String
import Foreign.C.Types
import GHC.TypeLits
import Numeric.LinearAlgebra.Static (R)
import qualified Numeric.LinearAlgebra.Static as H
import System.IO.Unsafe (unsafePerformIO)

-- ---------------------------------------------------------------------------------------------
-- C ABI (include/hamilton_b200.h)

data HbSystem
data HbOp = HbOp !Int32 !Int32 !Int32 !Double          -- op, a, b, c  (hb_op; 24 bytes with padding)

instance Storable HbOp where
  sizeOf _ = 24
  alignment _ = 8
  peek p = HbOp <$> peekByteOff p 0 <*> peekByteOff p 4 <*> peekByteOff p 8 <*> peekByteOff p 16
  poke p (HbOp o a b c) = pokeByteOff p 0 o >> pokeByteOff p 4 a >> pokeByteOff p 8 b >> pokeByteOff p 12 (0 :: Int32) >> pokeByteOff p 16 c

-- struct hb_tape { int32 n_in; int32 n_ops; const hb_op* ops; int32 n_out; const int32* outs; }
withTape :: Int -> [HbOp] -> [Int32] -> (Ptr () -> IO a) -> IO a
withTape nIn ops outs k =
  withArrayLen ops $ \nOps pOps -> withArrayLen outs $ \nOut pOut -> allocaBytes 32 $ \t -> do
    pokeByteOff t 0 (fromIntegral nIn :: Int32); pokeByteOff t 4 (fromIntegral nOps :: Int32)
    pokeByteOff t 8 pOps; pokeByteOff t 16 (fromIntegral nOut :: Int32); pokeByteOff t 24 pOut
    k t

foreign import ccall safe "hb_system_from_tape"      -- replaces mkSystem / mkSystem' (:201-254)
  c_system_from_tape :: Int32 -> Int32 -> Ptr Double -> Ptr () -> Ptr () -> Int32 -> Ptr Double -> Int32 -> Ptr (Ptr HbSystem) -> IO Int32
foreign import ccall safe "&hb_system_free" p_system_free :: FunPtr (Ptr HbSystem -> IO ())
foreign import ccall safe "hb_last_error" c_last_error :: IO CString
foreign import ccall safe "hb_underlying_pos" c_underlying_pos :: Ptr HbSystem -> Ptr Double -> Ptr Double -> IO Int32                     -- :174-178
foreign import ccall safe "hb_pe" c_pe :: Ptr HbSystem -> Ptr Double -> Ptr Double -> IO Int32                                              -- :182-186
foreign import ccall safe "hb_momenta" c_momenta :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32                     -- :262-269
foreign import ccall safe "hb_velocities" c_velocities :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32               -- :316-324
foreign import ccall safe "hb_ke_c" c_ke_c :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32                           -- :288-296
foreign import ccall safe "hb_ke_p" c_ke_p :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32                           -- :341-349
foreign import ccall safe "hb_lagrangian" c_lagrangian :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32               -- :301-309
foreign import ccall safe "hb_hamiltonian" c_hamiltonian :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32             -- :353-361
foreign import ccall safe "hb_ham_eqs" c_ham_eqs :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32       -- :370-387
foreign import ccall safe "hb_step_ham" c_step_ham :: Ptr HbSystem -> Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32   -- :390-402
foreign import ccall safe "hb_evolve_ham" c_evolve_ham :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> Int32 -> Ptr Double -> IO Int32 -- :433-462
foreign import ccall safe "hb_step_ham_c" c_step_ham_c :: Ptr HbSystem -> Double -> Ptr Double -> Ptr Double -> Ptr Double -> Ptr Double -> IO Int32 -- :505-515
foreign import ccall safe "hb_evolve_ham_c" c_evolve_ham_c :: Ptr HbSystem -> Ptr Double -> Ptr Double -> Ptr Double -> Int32 -> Ptr Double -> IO Int32 -- :488-498
foreign import ccall safe "hb_batch_step"
  c_batch_step :: Ptr HbSystem -> Int32 -> Double -> Int32 -> Int64 -> Int32 -> Int32 -> Ptr Double -> Ptr Double -> Ptr Int32 -> Ptr () -> IO Int32

-- | Non-zero status -> 'error', mirroring the reference's partiality (:425, :462, hmatrix's `inv`).
orDie :: String -> IO Int32 -> IO ()
orDie what act = do
  rc <- act
  when (rc /= 0) $ c_last_error >>= peekCString >>= \m -> error (what ++ ": " ++ m)

-- ---------------------------------------------------------------------------------------------
-- The tracing number type (the stand-in for ad's dual/tower numbers)

data Tr = Tr (IORef [HbOp]) Int32          -- tape (reversed) and node index

pushOp :: IORef [HbOp] -> HbOp -> Int32
pushOp ref o = unsafePerformIO $ atomicModifyIORef' ref (\os -> (o : os, fromIntegral (length os)))
{-# NOINLINE pushOp #-}

bin :: Int32 -> Tr -> Tr -> Tr
bin op (Tr r a) (Tr _ b) = Tr r (pushOp r (HbOp op a b 0))
un :: Int32 -> Tr -> Tr
un op (Tr r a) = Tr r (pushOp r (HbOp op a 0 0))
lit :: Tr -> Double -> Tr
lit (Tr r _) c = Tr r (pushOp r (HbOp 1 0 0 c))

-- hb_opcode numbering (include/hamilton_b200.h): ADD=3 SUB=4 MUL=5 DIV=6 NEG=7 RECIP=8 ABS=9 SIGNUM=10 SQRT=11 EXP=12
-- LOG=13 SIN=14 COS=15 TAN=16 ASIN=17 ACOS=18 ATAN=19 SINH=20 COSH=21 TANH=22 ASINH=23 ACOSH=24 ATANH=25 POW=26 ATAN2=28
-- NB: literals need a tape to live on; 'fromInteger'/'fromRational' create a detached constant that is re-homed by the
-- first binary operation it meets (elided here: see INTEGRATION.md "constants").
instance Num Tr where
  (+) = bin 3; (-) = bin 4; (*) = bin 5; negate = un 7; abs = un 9; signum = un 10
  fromInteger = error "re-homed constant (see INTEGRATION.md)"
instance Fractional Tr where
  (/) = bin 6; recip = un 8; fromRational = error "re-homed constant (see INTEGRATION.md)"
instance Floating Tr where
  pi = error "re-homed constant"; exp = un 12; log = un 13; sqrt = un 11; sin = un 14; cos = un 15; tan = un 16
  asin = un 17; acos = un 18; atan = un 19; sinh = un 20; cosh = un 21; tanh = un 22; asinh = un 23; acosh = un 24
  atanh = un 25; (**) = bin 26
-- RealFloat/RealFrac/Real/Ord/Eq instances: atan2 = bin 28; the comparison-based members raise, exactly the
-- data-dependent branching the tape cannot record (none of the reference's examples use it, app/Examples.hs).

-- ---------------------------------------------------------------------------------------------
-- Public API: same types as the reference

data Config (n :: Nat) = Cfg { cfgPositions :: !(R n), cfgVelocities :: !(R n) }      -- :103-113
data Phase (n :: Nat) = Phs { phsPositions :: !(R n), phsMomenta :: !(R n) }          -- :133-143
newtype System (m :: Nat) (n :: Nat) = Sys (ForeignPtr HbSystem)                       -- :160-169, abstract as in the reference

data Integrator = RK4 | RKF45_GSL deriving (Enum)

traceFn :: Int -> ([Tr] -> [Tr]) -> IO ([HbOp], [Int32])
traceFn nIn f = do
  ref <- newIORef []
  let ins = [Tr ref (pushOp ref (HbOp 0 (fromIntegral j) 0 0)) | j <- [0 .. nIn - 1]]
      outs = [i | Tr _ i <- f ins]
  mapM_ (\i -> i `seq` return ()) outs
  ops <- reverse <$> readIORef ref
  return (ops, outs)

mkSystemWith :: forall m n. (KnownNat m, KnownNat n) => Bool -> R m -> ([Tr] -> [Tr]) -> ([Tr] -> Tr) -> System m n
mkSystemWith onCart inertia f u = unsafePerformIO $ do
  let m = fromIntegral (natVal (Proxy @m)); n = fromIntegral (natVal (Proxy @n))
  (fops, fouts) <- traceFn n f
  (uops, uouts) <- traceFn (if onCart then m else n) (pure . u)
  withTape n fops fouts $ \ft -> withTape (if onCart then m else n) uops uouts $ \ut ->
    VS.unsafeWith (H.extract inertia) $ \pw -> alloca $ \out -> do
      orDie "mkSystem" $ c_system_from_tape (fromIntegral m) (fromIntegral n) pw ft ut (if onCart then 1 else 0) nullPtr 0 out
      Sys <$> (peek out >>= newForeignPtr p_system_free)
{-# NOINLINE mkSystemWith #-}

-- | src/Numeric/Hamilton.hs:201-225
mkSystem :: (KnownNat m, KnownNat n) => R m -> (forall a. RealFloat a => V.Vector n a -> V.Vector m a) -> (forall a. RealFloat a => V.Vector n a -> a) -> System m n
mkSystem = undefined   -- = mkSystemWith False, instantiating both arguments at Tr (needs the RealFloat Tr instance above)

-- | src/Numeric/Hamilton.hs:238-254
mkSystem' :: (KnownNat m, KnownNat n) => R m -> (forall a. RealFloat a => V.Vector n a -> V.Vector m a) -> (forall a. RealFloat a => V.Vector m a -> a) -> System m n
mkSystem' = undefined  -- = mkSystemWith True

withR :: KnownNat k => R k -> (Ptr Double -> IO a) -> IO a
withR v = VS.unsafeWith (H.extract v)

outR :: forall k. KnownNat k => (Ptr Double -> IO ()) -> IO (R k)
outR k = do
  let n = fromIntegral (natVal (Proxy @k))
  fp <- mallocForeignPtrArray n
  withForeignPtr fp k
  return (H.vector (VS.toList (VS.unsafeFromForeignPtr0 fp n)))

-- | hamEqs (:370-387)
hamEqs :: (KnownNat m, KnownNat n) => System m n -> Phase n -> (R n, R n)
hamEqs (Sys fp) (Phs q p) = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq -> withR p $ \pp -> do
  dqRef <- newIORef undefined
  dp <- outR $ \pdp -> do
    dq <- outR $ \pdq -> orDie "hamEqs" (c_ham_eqs s pq pp pdq pdp)
    writeIORef dqRef dq
  dq <- readIORef dqRef
  return (dq, dp)

-- | stepHam (:390-402): the library runs the same GSL-semantics adaptive RKF45 solve over (0, r).
stepHam :: (KnownNat m, KnownNat n) => Double -> System m n -> Phase n -> Phase n
stepHam r (Sys fp) (Phs q p) = unsafePerformIO $ withForeignPtr fp $ \s -> withR q $ \pq -> withR p $ \pp -> do
  poRef <- newIORef undefined
  qo <- outR $ \pqo -> do
    po <- outR $ \ppo -> orDie "stepHam" (c_step_ham s r pq pp pqo ppo)
    writeIORef poRef po
  Phs qo <$> readIORef poRef

-- The remaining wrappers (underlyingPos, pe, momenta, toPhase, keC, lagrangian, velocities, fromPhase, keP,
-- hamiltonian, evolveHam, evolveHam', stepHamC, evolveHamC, evolveHamC') follow the same two patterns and are
-- listed one-to-one in INTEGRATION.md.
underlyingPos = undefined; pe = undefined; momenta = undefined; toPhase = undefined; keC = undefined
lagrangian = undefined; velocities = undefined; fromPhase = undefined; keP = undefined; hamiltonian = undefined
evolveHam = undefined; evolveHam' = undefined; stepHamC = undefined; evolveHamC = undefined; evolveHamC' = undefined

-- | N trajectories at once: a storable vector of N Phases, laid out [q, p] per trajectory (HB_LAYOUT_AOS).
batchStep :: forall m n. (KnownNat m, KnownNat n) => Integrator -> Double -> Int -> System m n -> VS.Vector Double -> VS.Vector Double
batchStep integ dt nsteps (Sys fp) ys = unsafePerformIO $ withForeignPtr fp $ \s -> do
  let d = 2 * fromIntegral (natVal (Proxy @n)); nTraj = VS.length ys `div` d
  out <- mallocForeignPtrArray (VS.length ys)
  VS.unsafeWith ys $ \pin -> withForeignPtr out $ \pout ->
    orDie "batchStep" $ c_batch_step s (fromIntegral (fromEnum integ)) dt (fromIntegral nsteps) (fromIntegral nTraj) 0 0 pin pout nullPtr nullPtr
  return (VS.unsafeFromForeignPtr0 out (VS.length ys))
