-- | The range-proof layer: batched RangeProof.proveM / verifyM (src/RangeProof.hs:25-101) for TypedReciprocal and
-- Binary proofs behind bppp_rp_setup / bppp_rp_prove_batch / bppp_rp_verify_batch -- the API the benchmark uses.
-- Outputs are in the reference's order: coms = blCom : rCom : dmCom : mCom : nComs (TypedReciprocal.hs:441) or
-- blCom : dCom : nComs (Binary.hs:192), responses newest first (Bulletproof.hs:357-359), finals = getWitness
-- (norm scalars then linear scalars, RangeProof.hs:65): exactly the fields encodeProof' serialises (:60-66).
module RangeProof.B200
  ( Setup, Shape(..), Proof(..), newSetup, shape, useDeviceTranscript, proveBatch, verifyBatch ) where

import Control.Monad (forM)
import Foreign
import Foreign.C.String
import Foreign.C.Types

import Data.Curve (toA, fromA)
import Data.Curve.Weierstrass.SECP256K1 (PP, Fr)
import Data.Field.Galois (fromP, toP)

import Bulletproof.B200.FFI

newtype Setup = Setup (ForeignPtr RP)
data Shape = Shape { nInputs, numRpComs, nrmLen, linLen, rounds, finNorm, finLin :: Int } deriving Show
data Proof = Proof { coms :: [PP], responses :: [(PP, PP)], finals :: [Fr] }

-- | TRRP.setup (TypedReciprocal.hs:332-359) / setupBRP (Binary.hs:143-156): `binary`, argument kind (0 NL, 1 IP),
-- hasTypes / conserved, basisSeed of getPoints (app/Main.hs:68-72), the expanded ranges and public amounts
newSetup :: Bool -> Int -> Bool -> String -> [RangeSpec] -> [PublicSpec] -> IO Setup
newSetup binary arg flag seed ranges pubs =
  withCString seed $ \cseed -> withArrayLen ranges $ \nr pr -> withArrayLen pubs $ \np pp -> alloca $ \out -> do
    check "bppp_rp_setup" =<< c_rpSetup theCtx (b2i binary) (fromIntegral arg) (b2i flag) cseed 0 0
                                        (fromIntegral nr) pr (fromIntegral np) pp out
    Setup <$> (peek out >>= newForeignPtr p_rpFree)
  where b2i = fromIntegral . fromEnum

-- | infoRP + optimalWitnessSize: the shape every proof of this setup has
shape :: Setup -> IO Shape
shape (Setup fp) = withForeignPtr fp $ \s -> allocaArray 7 $ \v -> do
  let at i = v `advancePtr` i
  check "bppp_rp_info" =<< c_rpInfo s (at 0) (at 1) (at 2) (at 3) (at 4) (at 5) (at 6)
  [a, b, c, d, e, f, g] <- map fromIntegral <$> peekArray 7 v
  return (Shape a b c d e f g)

-- | run the Fiat-Shamir transcript of this setup's batches on the device (bit-identical proofs; SURVEY 8 f4)
useDeviceTranscript :: Setup -> Bool -> IO ()
useDeviceTranscript (Setup fp) on = withForeignPtr fp $ \s ->
  check "bppp_rp_set_device_transcript" =<< c_rpSetDeviceTranscript s (fromIntegral (fromEnum on))

-- | proveM for a batch: per proof its committed values, types and randomSeed (blinders derived like app/Main.hs:275-276)
proveBatch :: Setup -> [([Fr], [Fr], String)] -> IO [Proof]
proveBatch su@(Setup fp) inputs = do
  sh <- shape su
  let b = length inputs
      nc = numRpComs sh + nInputs sh
      nf = finNorm sh + finLin sh
      k = rounds sh
  seeds <- mapM (\(_, _, s) -> newCString s) inputs
  r <- withForeignPtr fp $ \s ->
    withLE32 (concatMap (\(vs, _, _) -> toInteger . fromP <$> vs) inputs) $ \pv ->
    withLE32 (concatMap (\(_, ts, _) -> toInteger . fromP <$> ts) inputs) $ \pt ->
    withArray seeds $ \pseeds ->
    allocaBytes (64 * b * nc) $ \pc -> allocaBytes (128 * b * k) $ \pr -> allocaBytes (32 * b * nf) $ \pf -> do
      check "bppp_rp_prove_batch" =<< c_rpProveBatch s (fromIntegral b) pv pt nullPtr pseeds pc pr pf
      forM [0 .. b - 1] $ \i -> do
        cs <- peekAffines64 nc (pc `plusPtr` (64 * i * nc))
        rs <- peekAffines64 (2 * k) (pr `plusPtr` (128 * i * k))
        fs <- mapM (\j -> toP <$> peekLE32 (pf `plusPtr` (32 * (i * nf + j)))) [0 .. nf - 1]
        return (Proof (fromA <$> cs) (pairs (fromA <$> rs)) fs)
  mapM_ free seeds
  return r
  where pairs (a : c : t) = (a, c) : pairs t
        pairs _ = []

-- | verifyM for a batch of proofs of this setup's shape
verifyBatch :: Setup -> [Proof] -> IO [Bool]
verifyBatch su@(Setup fp) proofs = do
  sh <- shape su
  let b = length proofs
  withForeignPtr fp $ \s ->
    withAffine64 (concatMap (map toA . coms) proofs) $ \pc ->
    withAffine64 (concatMap (concatMap (\(x, r) -> [toA x, toA r]) . responses) proofs) $ \pr ->
    withLE32 (concatMap (map (toInteger . fromP) . finals) proofs) $ \pf ->
    allocaArray b $ \pok -> do
      check "bppp_rp_verify_batch" =<< c_rpVerifyBatch s (fromIntegral b) (fromIntegral (rounds sh)) (fromIntegral (finNorm sh))
                                                       (fromIntegral (finLin sh)) pc pr pf pok
      map (/= 0) <$> peekArray b pok
