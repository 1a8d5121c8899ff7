{-# LANGUAGE ForeignFunctionInterface, EmptyDataDecls, ScopedTypeVariables #-}
-- | Raw bindings to libbppp_b200.so (include/bppp_b200.h) and the byte-level marshalling every seam shares.
--
-- Not compiled in the container this library is developed in (no GHC there); written against the
-- reference's package set (stack.yaml: lts-15.5 + elliptic-curve-0.3.0 + galois-field-1.0.1).  Add this
-- directory to `source-dirs` of the reference's package.yaml and `extra-libraries: [bppp_b200]`
-- (plus `extra-lib-dirs` pointing at bulletproofspp_b200/lib) to build it.
--
-- Conventions of the C ABI: a scalar / coordinate is a 32-byte little-endian canonical integer, an affine
-- point is x || y (64 bytes), the identity is 64 zero bytes; every call returns 0 or an error code and
-- never throws; results are pure functions of the inputs.
module Bulletproof.B200.FFI
  ( Ctx, Gens, NL, RP, Dtr
  , theCtx, check
  , withLE32, peekLE32, withAffine64, peekAffine64, peekAffines64
  , c_msm, c_msmBatch, c_pairFold, c_rationalReduce
  , c_gensCreate, c_gensMsmBatch
  , c_nlCreate, c_nlRoundCommit, c_nlRoundFold, c_nlFinal, c_nlVerify, p_nlDestroy
  , c_nlAttachTranscript, c_nlProveDevice
  , c_dtrCreate, c_dtrAbsorb, c_dtrSqueeze, c_dtrOracle, p_dtrDestroy
  , c_rpSetup, c_rpInfo, c_rpProveBatch, c_rpVerifyBatch, c_rpSetDeviceTranscript, p_rpFree
  , RangeSpec(..), PublicSpec(..)
  ) where

import Control.Monad (forM_, when, zipWithM_)
import Data.Bits (shiftL, shiftR, (.&.), (.|.))
import Data.IORef
import Data.Word
import Foreign
import Foreign.C.String
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)

import Data.Curve.Weierstrass (Point(A, O))
import Data.Curve.Weierstrass.SECP256K1 (PA, Fq, Fr)
import Data.Field.Galois (fromP, toP)

data Ctx
data Gens
data NL
data RP
data Dtr

foreign import ccall safe "bppp_init"       c_init      :: CInt -> Ptr (Ptr Ctx) -> IO CInt
foreign import ccall safe "bppp_last_error" c_lastError :: Ptr Ctx -> IO CString

-- ---- MSM seam (FastInnerProduct, src/Commitment.hs:311-353)
foreign import ccall safe "bppp_msm"
  c_msm :: Ptr Ctx -> CSize -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> IO CInt
foreign import ccall safe "bppp_msm_batch"
  c_msmBatch :: Ptr Ctx -> CSize -> CSize -> Ptr Word8 -> Ptr Word8 -> CInt -> Ptr Word8 -> IO CInt
foreign import ccall safe "bppp_pair_fold"
  c_pairFold :: Ptr Ctx -> CSize -> Ptr Word8 -> CInt -> Ptr Word8 -> CInt -> Ptr Word8 -> Ptr Word8 -> IO CInt
foreign import ccall unsafe "bppp_rational_reduce"
  c_rationalReduce :: Ptr Word8 -> Ptr Word8 -> Ptr CInt -> Ptr Word8 -> Ptr CInt -> IO CInt
foreign import ccall safe "bppp_gens_create"
  c_gensCreate :: Ptr Ctx -> CSize -> CSize -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr (Ptr Gens) -> IO CInt
foreign import ccall safe "bppp_gens_msm_batch"
  c_gensMsmBatch :: Ptr Gens -> CSize -> CSize -> Ptr Word8 -> Ptr Word8 -> IO CInt

-- ---- argument seam (BPOpening of NL.NormLinear behind proveRoundM / verifyBPM, src/Bulletproof.hs:276-291, 346-378)
foreign import ccall safe "bppp_nl_create"
  c_nlCreate :: Ptr Ctx -> CInt -> CSize -> CSize -> CSize
             -> Ptr Word8 -> Ptr Word8 -> Ptr Word8                                  -- g, G, H
             -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr Word8        -- q, s, w, l, c
             -> Ptr (Ptr NL) -> IO CInt
foreign import ccall safe "bppp_nl_round_commit" c_nlRoundCommit :: Ptr NL -> Ptr Word8 -> Ptr Word8 -> IO CInt
foreign import ccall safe "bppp_nl_round_fold"   c_nlRoundFold   :: Ptr NL -> Ptr Word8 -> IO CInt
foreign import ccall safe "bppp_nl_final"        c_nlFinal       :: Ptr NL -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> IO CInt
foreign import ccall safe "&bppp_nl_destroy"     p_nlDestroy     :: FunPtr (Ptr NL -> IO ())
foreign import ccall safe "bppp_nl_verify"
  c_nlVerify :: Ptr Ctx -> CInt -> CSize -> CSize -> CSize -> CSize
             -> Ptr Word8 -> Ptr Word8 -> Ptr Word8                                  -- g, G, H
             -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr Word8                     -- q, s_pub, pub_w, c
             -> Ptr Word8 -> Ptr Word8 -> CSize -> CSize -> Ptr Word8 -> Ptr Word8   -- es, XR, n_norm, n_lin, fw, fl
             -> CSize -> Ptr Word8 -> Ptr Word8 -> Ptr CInt -> IO CInt               -- n_init, init_s, init_p, ok
-- the whole round loop on the device, transcript included (SURVEY 8 f4)
foreign import ccall safe "bppp_nl_attach_transcript" c_nlAttachTranscript :: Ptr NL -> Ptr Dtr -> IO CInt
foreign import ccall safe "bppp_nl_prove_device"
  c_nlProveDevice :: Ptr NL -> CSize -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> IO CInt

-- ---- the reference's transcript on the device (shaOracle, app/Main.hs:75-80; ZKPT, src/ZKP.hs:68-101)
foreign import ccall safe "bppp_dtr_create"   c_dtrCreate  :: Ptr Ctx -> CSize -> CSize -> CInt -> Ptr (Ptr Dtr) -> IO CInt
foreign import ccall safe "bppp_dtr_absorb"   c_dtrAbsorb  :: Ptr Dtr -> Ptr Word8 -> CSize -> CSize -> IO CInt
foreign import ccall safe "bppp_dtr_squeeze"  c_dtrSqueeze :: Ptr Dtr -> CSize -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> IO CInt
foreign import ccall safe "bppp_dtr_oracle"   c_dtrOracle  :: Ptr Dtr -> Ptr Word8 -> CSize -> CInt -> Ptr Word8 -> IO CInt
foreign import ccall safe "&bppp_dtr_destroy" p_dtrDestroy :: FunPtr (Ptr Dtr -> IO ())

-- ---- range-proof layer (RPOpening / RangeProof.proveM / verifyM, src/RangeProof.hs:25-101)
-- | post-`count`-expansion RangeData of app/Parse.hs:126-172; min / max are 128-bit two's complement, little-endian
data RangeSpec = RangeSpec { rsMin :: Integer, rsMax :: Integer, rsBase :: Word32
                           , rsShared :: Bool, rsOutput :: Bool, rsAssumed :: Bool }
-- | PubSpec of app/Parse.hs:210-235
data PublicSpec = PublicSpec { psAmount :: Integer, psType :: Integer, psOutput :: Bool }

pokeI128 :: Ptr Word8 -> Integer -> IO ()
pokeI128 p v = forM_ [0 .. 15] $ \i -> pokeByteOff p i (fromIntegral ((v `shiftR` (8 * i)) .&. 255) :: Word8)

instance Storable RangeSpec where            -- struct bppp_range_spec: u8 min[16], u8 max[16], u32 base, 3 x i32
  sizeOf _ = 48
  alignment _ = 4
  poke p (RangeSpec mn mx b sh o a) = do
    pokeI128 (castPtr p) mn
    pokeI128 (castPtr p `plusPtr` 16) mx
    pokeByteOff p 32 b
    pokeByteOff p 36 (fromIntegral (fromEnum sh) :: CInt)
    pokeByteOff p 40 (fromIntegral (fromEnum o) :: CInt)
    pokeByteOff p 44 (fromIntegral (fromEnum a) :: CInt)
  peek _ = error "RangeSpec is write-only"
instance Storable PublicSpec where           -- struct bppp_public_spec: u8 amount[16], u8 type[16], i32 is_output
  sizeOf _ = 36
  alignment _ = 4
  poke p (PublicSpec am ty o) = do
    pokeI128 (castPtr p) am
    pokeI128 (castPtr p `plusPtr` 16) ty
    pokeByteOff p 32 (fromIntegral (fromEnum o) :: CInt)
  peek _ = error "PublicSpec is write-only"

foreign import ccall safe "bppp_rp_setup"
  c_rpSetup :: Ptr Ctx -> CInt -> CInt -> CInt -> CString -> CInt -> CInt
            -> CSize -> Ptr RangeSpec -> CSize -> Ptr PublicSpec -> Ptr (Ptr RP) -> IO CInt
foreign import ccall safe "bppp_rp_info"
  c_rpInfo :: Ptr RP -> Ptr CSize -> Ptr CSize -> Ptr CSize -> Ptr CSize -> Ptr CSize -> Ptr CSize -> Ptr CSize -> IO CInt
foreign import ccall safe "bppp_rp_prove_batch"
  c_rpProveBatch :: Ptr RP -> CSize -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr CString
                 -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> IO CInt
foreign import ccall safe "bppp_rp_verify_batch"
  c_rpVerifyBatch :: Ptr RP -> CSize -> CSize -> CSize -> CSize -> Ptr Word8 -> Ptr Word8 -> Ptr Word8 -> Ptr CInt -> IO CInt
foreign import ccall safe "bppp_rp_set_device_transcript" c_rpSetDeviceTranscript :: Ptr RP -> CInt -> IO CInt
foreign import ccall safe "&bppp_rp_free" p_rpFree :: FunPtr (Ptr RP -> IO ())

-- ---- one process-wide context on device 0 (bppp_init creates a stream and sizes the memory pool)
{-# NOINLINE theCtx #-}
theCtx :: Ptr Ctx
theCtx = unsafePerformIO $ alloca $ \pp -> do
  rc <- c_init 0 pp
  when (rc /= 0) $ error ("bppp_init failed: " ++ show rc ++ " (no CUDA device?  there is no CPU fallback)")
  peek pp

-- | 0 or die with the library's message
check :: String -> CInt -> IO ()
check what rc = when (rc /= 0) $ do
  msg <- c_lastError theCtx >>= peekCString
  error (what ++ " failed (" ++ show rc ++ "): " ++ msg)

-- ---- marshalling
pokeLE32 :: Ptr Word8 -> Integer -> IO ()
pokeLE32 p v = forM_ [0 .. 31] $ \i -> pokeByteOff p i (fromIntegral ((v `shiftR` (8 * i)) .&. 255) :: Word8)

peekLE32 :: Ptr Word8 -> IO Integer
peekLE32 p = foldr (\b acc -> (acc `shiftL` 8) .|. fromIntegral (b :: Word8)) 0 <$> peekArray 32 p

-- | canonical representatives (`fromP`) of field elements as consecutive 32-byte little-endian integers
withLE32 :: [Integer] -> (Ptr Word8 -> IO a) -> IO a
withLE32 xs k = allocaBytes (32 * max 1 (length xs)) $ \p -> do
  zipWithM_ (\i x -> pokeLE32 (p `plusPtr` (32 * i)) x) [0 ..] xs
  k p

-- | affine points as consecutive x || y records, the identity as 64 zero bytes
withAffine64 :: [PA] -> (Ptr Word8 -> IO a) -> IO a
withAffine64 ps k = allocaBytes (64 * max 1 (length ps)) $ \p -> do
  zipWithM_ (\i pt -> pokePoint (p `plusPtr` (64 * i)) pt) [0 ..] ps
  k p
  where
    pokePoint q O = pokeLE32 q 0 >> pokeLE32 (q `plusPtr` 32) 0
    pokePoint q (A x y) = pokeLE32 q (toInteger (fromP x)) >> pokeLE32 (q `plusPtr` 32) (toInteger (fromP y))

peekAffine64 :: Ptr Word8 -> IO PA
peekAffine64 p = do
  x <- peekLE32 p
  y <- peekLE32 (p `plusPtr` 32)
  return $ if x == 0 && y == 0 then O else A (toP x :: Fq) (toP y :: Fq)

peekAffines64 :: Int -> Ptr Word8 -> IO [PA]
peekAffines64 n p = mapM (\i -> peekAffine64 (p `plusPtr` (64 * i))) [0 .. n - 1]
