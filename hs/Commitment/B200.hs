{-# LANGUAGE TypeFamilies, GeneralizedNewtypeDeriving, FlexibleInstances, ScopedTypeVariables #-}
-- | The MSM seam: a `FastInnerProduct` instance whose `innerProduct` and `projectivePairIP` run on the B200.
--
-- Replaces the 256-row Straus loop of src/Commitment.hs:325-335 and the 129-row pair loop of :343-353 for every
-- `commit` (src/Commitment.hs:416-417) and every `collapsePoints` fold (src/Bulletproof.hs:213-214) with zero
-- changes above Commitment.hs: switch `type PX = PP` to `type PX = GpuPoint` in app/Main.hs:105 ("Change this to
-- change coordinates").  Batch size 1 per call -- the compatibility seam; throughput comes from the argument
-- and range-proof seams (Bulletproof.B200, RangeProof.B200).
module Commitment.B200 (GpuPoint(..)) where

import Foreign (allocaBytes)
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)
import Control.DeepSeq (NFData)

import Data.Curve (toA, fromA)
import Data.Curve.Weierstrass.SECP256K1 (PP, PA, Fr)
import Data.Field.Galois (fromP)
import Data.VectorSpace

import Commitment
import Bulletproof.B200.FFI

-- | secp256k1 point in the reference's projective coordinates; the group structure is the wrapped type's
newtype GpuPoint = GP { unGP :: PP } deriving (Eq, Show, NFData)

instance AdditiveGroup GpuPoint where
  zeroV = GP zeroV
  GP a ^+^ GP b = GP (a ^+^ b)
  negateV (GP a) = GP (negateV a)
instance VectorSpace GpuPoint where
  type Scalar GpuPoint = Fr
  s *^ GP a = GP (s *^ a)
instance FastDouble GpuPoint where
  dbl' (GP a) = GP (dbl' a)

affine :: GpuPoint -> PA
affine = toA . unGP

instance FastInnerProduct GpuPoint where
  -- the device keeps no per-call basis: the class's row-wise helpers are never reached
  type Basis GpuPoint = GpuPoint
  addBasis _ _ v = v
  normalizeBasis = id

  -- innerProduct :: [(Scalar v, v)] -> v     (src/Commitment.hs:325-335)  ->  bppp_msm
  innerProduct [] = zeroV
  innerProduct sgs = unsafePerformIO $
    withLE32 (toInteger . fromP . fst <$> sgs) $ \ps ->
    withAffine64 (affine . snd <$> sgs) $ \pp ->
    allocaBytes 64 $ \out -> do
      check "bppp_msm" =<< c_msm theCtx (fromIntegral (length sgs)) ps pp out
      GP . fromA <$> peekAffine64 out

  -- projectivePairIP (s0, g0) (s1, g1) = s0 g0 + s1 g1 for the short pair of rationalReduceScalar
  -- (src/Commitment.hs:343-353; reduced scalars are plain Integers for `Prime p`, :269-288)  ->  bppp_pair_fold
  projectivePairIP (b, gL) (a, gR) = unsafePerformIO $
    withLE32 [abs a] $ \pa -> withLE32 [abs b] $ \pb ->
    withAffine64 [affine gL, affine gR] $ \pin ->
    allocaBytes 64 $ \out -> do
      check "bppp_pair_fold" =<< c_pairFold theCtx 2 pa (sgn a) pb (sgn b) pin out
      GP . fromA <$> peekAffine64 out
    where sgn v = if v < 0 then 1 else 0 :: CInt
