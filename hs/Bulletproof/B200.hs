{-# LANGUAGE ScopedTypeVariables #-}
-- | The argument seam: proveBPM / verifyBPM of the norm-linear argument (src/Bulletproof.hs:346-378 over
-- NL.NormLinear, src/Bulletproof/NormArgument.hs) with the vectors resident on the B200.
--
-- `proveBPMDevice` keeps the reference's round structure -- it is `proveBPM` with `makeScalarsComs` + the two
-- `commit`s of `proveRoundM` replaced by `bppp_nl_round_commit` and `collapse e` by `bppp_nl_round_fold`; the
-- challenge still comes from the caller's own `oracle [ac, bc]`, i.e. the ZKPT transcript on the host.
-- `proveBPMOnDevice` is the variant with the transcript on the device (SURVEY 8 f4): the whole loop is one call.
module Bulletproof.B200
  ( DeviceArg, newDeviceArg, proveBPMDevice, proveBPMOnDevice, verifyBPMDevice ) where

import Control.Monad (foldM)
import Foreign
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)

import Data.Curve (toA, fromA)
import Data.Curve.Weierstrass.SECP256K1 (PP, PA, Fr)
import Data.Field.Galois (fromP, toP)

import ZKP (oracle, ZKP)
import Bulletproof.B200.FFI

newtype DeviceArg = DeviceArg (ForeignPtr NL)

fr :: Fr -> Integer
fr = toInteger . fromP

-- | makeNormLinearBP' (src/Bulletproof.hs:283-285) for one proof: scalar weight 1, weight q, linear coefficients
-- cs, norm witness over gs, linear witness over hs, scalar part s on g.
newDeviceArg :: PP -> [PP] -> [PP] -> Fr -> Fr -> [Fr] -> [Fr] -> [Fr] -> IO DeviceArg
newDeviceArg g gs hs q s ws ls cs =
  withAffine64 [toA g] $ \pg -> withAffine64 (toA <$> gs) $ \pG -> withAffine64 (toA <$> hs) $ \pH ->
  withLE32 [fr q] $ \pq -> withLE32 [fr s] $ \ps -> withLE32 (fr <$> ws) $ \pw ->
  withLE32 (fr <$> ls) $ \pl -> withLE32 (fr <$> cs) $ \pc -> alloca $ \out -> do
    check "bppp_nl_create" =<< c_nlCreate theCtx 0 1 (fromIntegral (length gs)) (fromIntegral (length hs))
                                          pg pG pH pq ps pw pl pc out
    DeviceArg <$> (peek out >>= newForeignPtr p_nlDestroy)

-- | proveBPM n (src/Bulletproof.hs:357-359): responses newest first, then (s, norm witness, linear witness)
proveBPMDevice :: ZKP PP Fr m => Int -> (Int, Int) -> DeviceArg -> m ([(PP, PP)], (Fr, [Fr], [Fr]))
proveBPMDevice n (finN, finL) (DeviceArg fp) = do
  resps <- foldM step [] [1 .. n]
  return (resps, final)
  where
    step acc _ = do
      let (x, r) = unsafeRound
      e <- head <$> oracle [x, r]                      -- e <- head <$> oracle [ac, bc]   (Bulletproof.hs:351)
      unsafeFold e `seq` return ((x, r) : acc)
    unsafeRound = unsafeIO $ withForeignPtr fp $ \h -> allocaBytes 64 $ \px -> allocaBytes 64 $ \pr -> do
      check "bppp_nl_round_commit" =<< c_nlRoundCommit h px pr
      (,) <$> (fromA <$> peekAffine64 px) <*> (fromA <$> peekAffine64 pr)
    unsafeFold e = unsafeIO $ withForeignPtr fp $ \h -> withLE32 [fr e] $ \pe ->
      check "bppp_nl_round_fold" =<< c_nlRoundFold h pe
    final = unsafeIO $ withForeignPtr fp $ \h ->
      allocaBytes 32 $ \ps -> allocaBytes (32 * max 1 finN) $ \pw -> allocaBytes (32 * max 1 finL) $ \pl -> do
        check "bppp_nl_final" =<< c_nlFinal h ps pw pl
        s <- toP <$> peekLE32 ps
        ws <- mapM (\i -> toP <$> peekLE32 (pw `plusPtr` (32 * i))) [0 .. finN - 1]
        ls <- mapM (\i -> toP <$> peekLE32 (pl `plusPtr` (32 * i))) [0 .. finL - 1]
        return (s, ws, ls)
    -- the handle is mutated in place; the data dependencies (x, r) -> e -> fold -> next (x, r) order the calls
    unsafeIO :: IO a -> a
    unsafeIO = unsafePerformIO

-- | The same proof with the Fiat-Shamir transcript on the device: `initial` are the commitments already in the
-- transcript, newest first (for a bare argument: its initial commitment).  One library call, one synchronisation.
proveBPMOnDevice :: Int -> (Int, Int) -> [PP] -> DeviceArg -> IO ([(PP, PP)], [Fr], (Fr, [Fr], [Fr]))
proveBPMOnDevice n (finN, finL) initial (DeviceArg fp) =
  withForeignPtr fp $ \h -> alloca $ \pt -> do
    check "bppp_dtr_create" =<< c_dtrCreate theCtx 1 (fromIntegral (length initial + 2 * n)) 0 pt
    t <- peek pt >>= newForeignPtr p_dtrDestroy
    withForeignPtr t $ \tr -> do
      withAffine64 (toA <$> initial) $ \pin ->
        check "bppp_dtr_absorb" =<< c_dtrAbsorb tr pin (fromIntegral (length initial)) (fromIntegral (length initial))
      check "bppp_nl_attach_transcript" =<< c_nlAttachTranscript h tr
      allocaBytes (128 * n) $ \presp -> allocaBytes (32 * n) $ \pes ->
        allocaBytes 32 $ \ps -> allocaBytes (32 * max 1 finN) $ \pw -> allocaBytes (32 * max 1 finL) $ \pl -> do
          check "bppp_nl_prove_device" =<< c_nlProveDevice h (fromIntegral n) presp pes ps pw pl
          pts <- peekAffines64 (2 * n) presp
          es <- mapM (\i -> toP <$> peekLE32 (pes `plusPtr` (32 * i))) [0 .. n - 1]
          s <- toP <$> peekLE32 ps
          ws <- mapM (\i -> toP <$> peekLE32 (pw `plusPtr` (32 * i))) [0 .. finN - 1]
          ls <- mapM (\i -> toP <$> peekLE32 (pl `plusPtr` (32 * i))) [0 .. finL - 1]
          return (pairs (fromA <$> pts), es, (s, ws, ls))
  where pairs (a : b : r) = (a, b) : pairs r
        pairs _ = []

-- | verifyBPM's collapsed check (src/Bulletproof.hs:370-378) for one proof; `es` newest first like the responses
verifyBPMDevice :: PP -> [PP] -> [PP] -> Fr -> Fr -> [Fr] -> [Fr] -> [Fr] -> [(PP, PP)] -> [Fr] -> [Fr]
                -> [(Fr, PP)] -> IO Bool
verifyBPMDevice g gs hs q sPub pubW cs es resps fw fl initial =
  withAffine64 [toA g] $ \pg -> withAffine64 (toA <$> gs) $ \pG -> withAffine64 (toA <$> hs) $ \pH ->
  withLE32 [fr q] $ \pq -> withLE32 [fr sPub] $ \psp -> withLE32 (fr <$> pubW) $ \ppw -> withLE32 (fr <$> cs) $ \pc ->
  withLE32 (fr <$> es) $ \pes -> withAffine64 (concatMap (\(x, r) -> [toA x, toA r]) resps) $ \pxr ->
  withLE32 (fr <$> fw) $ \pfw -> withLE32 (fr <$> fl) $ \pfl ->
  withLE32 (fr . fst <$> initial) $ \pis -> withAffine64 (toA . snd <$> initial) $ \pip -> alloca $ \pok -> do
    check "bppp_nl_verify" =<< c_nlVerify theCtx 0 1 (fromIntegral (length gs)) (fromIntegral (length hs))
                                          (fromIntegral (length resps)) pg pG pH pq psp ppw pc pes pxr
                                          (fromIntegral (length fw)) (fromIntegral (length fl)) pfw pfl
                                          (fromIntegral (length initial)) pis pip pok
    (/= 0) <$> peek pok
