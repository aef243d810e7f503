//! Drop-in batch forms of the issuer side of `anonymous-credit-tokens` backed by the B200 engine.
//!
//! The reference's protocol structs have private fields, so records cross the boundary as the crate's own CBOR
//! (`to_cbor` / `from_cbor`, src/cbor.rs): canonical items are unpacked on the device (`act_unpack_cbor`), anything else
//! by the lenient host parser (`act_pack_*_cbor`) -- same accept/reject rules as `from_cbor`.
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image); see ../README.md.
use act_b200_sys as sys;
use anonymous_credit_tokens::{Error, IssuanceRequest, IssuanceResponse, Params, PrivateKey, Refund, SpendProof};
use curve25519_dalek::Scalar;
use rand_core::CryptoRngCore;

pub struct Engine {
    raw: *mut sys::act_engine,
}
unsafe impl Send for Engine {}

fn status_to_error(st: u8) -> Error {
    match st {
        1 => Error::InvalidIssuanceRequestProof,
        2 => Error::InvalidIssuanceResponseProof,
        3 => Error::DoubleSpendError,
        4 => Error::InvalidRefundProof,
        5 => Error::InvalidRefundResponseProof,
        6 => Error::IdentityPointError,
        7 => Error::InvalidClientSpendProof,
        8 => Error::AmountTooBigError,
        _ => Error::ScalarOutOfRangeError,
    }
}

impl Engine {
    /// `domain` = the four strings given to `Params::new` (the reference has no accessor or CBOR form for `Params`,
    /// so H1..H3 are re-derived on the device: src/lib.rs:291-354); the key crosses as its CBOR (`{1: x, 2: W}`, src/cbor.rs:476-485).
    pub fn new(domain: [&str; 4], key: &PrivateKey, device: i32) -> Result<Self, String> {
        let c: Vec<std::ffi::CString> = domain.iter().map(|s| std::ffi::CString::new(*s).unwrap()).collect();
        let mut h = [0u8; 96];
        if unsafe { sys::act_params_derive(device, c[0].as_ptr(), c[1].as_ptr(), c[2].as_ptr(), c[3].as_ptr(), h.as_mut_ptr()) } != 0 {
            return Err(last_error());
        }
        let kc = key.to_cbor().map_err(|e| format!("{e:?}"))?;
        // canonical layout: a2 01 58 20 <x:32> 02 58 20 <W:32>
        if kc.len() != 71 || kc[0] != 0xa2 {
            return Err("unexpected PrivateKey CBOR layout".into());
        }
        let (x, w) = (&kc[4..36], &kc[39..71]);
        let mut raw = std::ptr::null_mut();
        if unsafe { sys::act_engine_create(&mut raw, device, h.as_ptr(), x.as_ptr(), w.as_ptr()) } != 0 {
            return Err(last_error());
        }
        Ok(Engine { raw })
    }

    /// One handle over several GPUs of the box (`act_engine_create_multi`): every call below shards the slice contiguously
    /// over one replica per device; results are identical to the single-device engine's.
    pub fn new_multi(domain: [&str; 4], key: &PrivateKey, devices: &[i32]) -> Result<Self, String> {
        let c: Vec<std::ffi::CString> = domain.iter().map(|s| std::ffi::CString::new(*s).unwrap()).collect();
        let mut h = [0u8; 96];
        if unsafe { sys::act_params_derive(devices[0], c[0].as_ptr(), c[1].as_ptr(), c[2].as_ptr(), c[3].as_ptr(), h.as_mut_ptr()) } != 0 {
            return Err(last_error());
        }
        let kc = key.to_cbor().map_err(|e| format!("{e:?}"))?;
        if kc.len() != 71 || kc[0] != 0xa2 {
            return Err("unexpected PrivateKey CBOR layout".into());
        }
        let (x, w) = (&kc[4..36], &kc[39..71]);
        let mut raw = std::ptr::null_mut();
        if unsafe { sys::act_engine_create_multi(&mut raw, devices.as_ptr(), devices.len() as i32, h.as_ptr(), x.as_ptr(), w.as_ptr()) } != 0 {
            return Err(last_error());
        }
        Ok(Engine { raw })
    }

    /// CBOR items -> fixed records: the canonical skeleton on the device, everything else through the host parser.
    fn unpack(&self, kind: i32, items: &[Vec<u8>], cbor_len: usize, rec_len: usize) -> (Vec<u8>, Vec<u8>) {
        let n = items.len();
        let mut rec = vec![0u8; n * rec_len];
        let mut st = vec![sys::ACT_STATUS_NOT_CANONICAL; n];
        let fixed: Vec<usize> = (0..n).filter(|&i| items[i].len() == cbor_len).collect();
        if !fixed.is_empty() {
            let mut flat = Vec::with_capacity(fixed.len() * cbor_len);
            for &i in &fixed { flat.extend_from_slice(&items[i]); }
            let mut r = vec![0u8; fixed.len() * rec_len];
            let mut s = vec![0u8; fixed.len()];
            let rc = unsafe { sys::act_unpack_cbor(self.raw, kind, fixed.len(), flat.as_ptr(), r.as_mut_ptr(), s.as_mut_ptr()) };
            assert_eq!(rc, 0, "{}", last_error());
            for (k, &i) in fixed.iter().enumerate() {
                st[i] = s[k];
                rec[i * rec_len..(i + 1) * rec_len].copy_from_slice(&r[k * rec_len..(k + 1) * rec_len]);
            }
        }
        for i in 0..n {
            if st[i] == sys::ACT_STATUS_NOT_CANONICAL {
                let p = items[i].as_ptr();
                let l = items[i].len();
                let f = match kind {
                    sys::ACT_KIND_REQUEST => sys::act_pack_issuance_requests_cbor,
                    sys::ACT_KIND_PROOF => sys::act_pack_spend_proofs_cbor,
                    sys::ACT_KIND_RESPONSE => sys::act_pack_issuance_responses_cbor,
                    _ => sys::act_pack_refunds_cbor,
                };
                unsafe { f(1, &p, &l, rec[i * rec_len..].as_mut_ptr(), st[i..].as_mut_ptr()) };
            }
        }
        (rec, st)
    }

    /// Batch form of `PrivateKey::issue` (src/lib.rs:621-663) with the semantics of a loop over ONE shared RNG: same outputs, and
    /// the same RNG state afterwards (128 bytes drawn per accepted request, none for a rejected one).
    pub fn batch_issue(&self, reqs: &[IssuanceRequest], cs: &[Scalar], mut rng: impl CryptoRngCore)
        -> Vec<Result<IssuanceResponse, Error>> {
        let n = reqs.len();
        assert_eq!(cs.len(), n);
        let items: Vec<Vec<u8>> = reqs.iter().map(|r| r.to_cbor().expect("to_cbor")).collect();
        let (rec, pst) = self.unpack(sys::ACT_KIND_REQUEST, &items, sys::ACT_CBOR_REQUEST_BYTES, sys::ACT_REQUEST_BYTES);
        debug_assert!(pst.iter().all(|&s| s == 0));
        let mut c = vec![0u8; 32 * n];
        for (i, s) in cs.iter().enumerate() { c[32 * i..32 * i + 32].copy_from_slice(s.as_bytes()); }
        // Verify first, then draw: the reference takes e and alpha from the RNG only after a request verifies
        // (src/lib.rs:638-643), so a loop of issue() calls leaves the caller's RNG advanced by 128 bytes per ACCEPTED request.
        // Drawing exactly that many bytes, in slice order, leaves `rng` in the same state as the loop would.
        let (mut resp, mut st) = (vec![0u8; sys::ACT_RESPONSE_BYTES * n], vec![0u8; n]);
        let rc = unsafe { sys::act_batch_issue_verify(self.raw, n, rec.as_ptr(), st.as_mut_ptr()) };
        assert_eq!(rc, 0, "{}", last_error());
        let accepted = st.iter().filter(|&&s| s == 0).count();
        let mut stream = vec![0u8; 128 * accepted];
        for chunk in stream.chunks_mut(64) { rng.fill_bytes(chunk); }   // one 64-byte draw per Scalar::random, as the reference does
        let rc = unsafe { sys::act_batch_issue_sign(self.raw, n, rec.as_ptr(), c.as_ptr(), st.as_ptr(), stream.as_ptr(), stream.len(), resp.as_mut_ptr()) };
        assert_eq!(rc, 0, "{}", last_error());
        stream.iter_mut().for_each(|b| *b = 0);
        let mut cb = vec![0u8; sys::ACT_CBOR_RESPONSE_BYTES * n];
        unsafe { sys::act_encode_cbor(self.raw, sys::ACT_KIND_RESPONSE, n, resp.as_ptr(), cb.as_mut_ptr()) };
        (0..n).map(|i| if st[i] == 0 {
            Ok(IssuanceResponse::from_cbor(&cb[sys::ACT_CBOR_RESPONSE_BYTES * i..sys::ACT_CBOR_RESPONSE_BYTES * (i + 1)]).expect("engine output"))
        } else { Err(status_to_error(st[i])) }).collect()
    }

    /// Batch form of `PrivateKey::refund` (src/lib.rs:781-869) returning `SpendProof::nullifier()` with each refund.
    pub fn batch_verify_spend_and_refund(&self, proofs: &[SpendProof], mut rng: impl CryptoRngCore)
        -> Vec<Result<(Scalar, Refund), Error>> {
        let n = proofs.len();
        let items: Vec<Vec<u8>> = proofs.iter().map(|p| p.to_cbor().expect("to_cbor")).collect();
        let (rec, pst) = self.unpack(sys::ACT_KIND_PROOF, &items, sys::ACT_CBOR_PROOF_BYTES, sys::ACT_PROOF_BYTES);
        debug_assert!(pst.iter().all(|&s| s == 0));
        // verify, count, draw 128 bytes per accepted proof in slice order (src/lib.rs:842-846), sign: see batch_issue
        let (mut refunds, mut nul, mut st, mut kprime) = (vec![0u8; 128 * n], vec![0u8; 32 * n], vec![0u8; n], vec![0u8; 128 * n]);
        let rc = unsafe { sys::act_batch_spend_verify(self.raw, n, rec.as_ptr(), nul.as_mut_ptr(), st.as_mut_ptr(), kprime.as_mut_ptr()) };
        assert_eq!(rc, 0, "{}", last_error());
        let accepted = st.iter().filter(|&&s| s == 0).count();
        let mut stream = vec![0u8; 128 * accepted];
        for chunk in stream.chunks_mut(64) { rng.fill_bytes(chunk); }
        let rc = unsafe { sys::act_batch_refund_sign(self.raw, n, kprime.as_ptr(), st.as_ptr(), stream.as_ptr(), stream.len(), refunds.as_mut_ptr()) };
        assert_eq!(rc, 0, "{}", last_error());
        stream.iter_mut().for_each(|b| *b = 0);
        let mut cb = vec![0u8; sys::ACT_CBOR_REFUND_BYTES * n];
        unsafe { sys::act_encode_cbor(self.raw, sys::ACT_KIND_REFUND, n, refunds.as_ptr(), cb.as_mut_ptr()) };
        (0..n).map(|i| if st[i] == 0 {
            let mut k = [0u8; 32];
            k.copy_from_slice(&nul[32 * i..32 * i + 32]);
            let r = Refund::from_cbor(&cb[sys::ACT_CBOR_REFUND_BYTES * i..sys::ACT_CBOR_REFUND_BYTES * (i + 1)]).expect("engine output");
            Ok((Scalar::from_bytes_mod_order(k), r))
        } else { Err(status_to_error(st[i])) }).collect()
    }

    /// The caller's nullifier check over a batch (src/lib.rs:741-745): later duplicates and members of `seen` -> DoubleSpendError.
    pub fn flag_replays(&self, status: &[u8], nullifiers: &[u8], seen: &[u8]) -> Vec<u8> {
        let mut out = vec![0u8; status.len()];
        let rc = unsafe { sys::act_flag_replays(self.raw, status.len(), status.as_ptr(), nullifiers.as_ptr(), seen.len() / 32,
                                                if seen.is_empty() { std::ptr::null() } else { seen.as_ptr() }, out.as_mut_ptr()) };
        assert_eq!(rc, 0, "{}", last_error());
        out
    }
}

impl Drop for Engine {
    fn drop(&mut self) { unsafe { sys::act_engine_destroy(self.raw) } }
}

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(sys::act_last_error()).to_string_lossy().into_owned() }
}

#[allow(dead_code)]
fn _params_is_only_a_witness(_: &Params) {}
