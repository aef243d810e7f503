//! Known-answer vectors from the REAL `anonymous-credit-tokens` crate, for `tests/golden/ref_crate.json`.
//!
//! The crate's own tests hold no known-answer values (every test draws from `OsRng`, src/tests.rs:17), so this
//! repository's parity is pinned only indirectly.  This program closes that gap wherever a Rust
//! toolchain exists:
//!
//!     cd rust/golden-dump && cargo run --release > ../../tests/golden/ref_crate.json
//!
//! It drives the crate's public API (src/lib.rs:188,432,463,528,621,781,972,1217) from a ChaCha20 stream with a fixed
//! seed, through an RNG wrapper that RECORDS every byte each call draws.  The JSON therefore carries, for every call,
//! the inputs as the crate's own CBOR (`to_cbor`, src/cbor.rs), the exact RNG bytes the call consumed, and the outcome
//! (output CBOR, or the `Error` variant) -- so the consumer needs no ChaCha: `tests/test_ref_crate_golden.py` replays
//! the recorded bytes through the CPU oracle and through the CUDA engine and compares byte for byte.
//!
//! Mutations are made on CBOR bytes and re-parsed with the crate's own `from_cbor` (the protocol structs' fields are
//! private, src/lib.rs:376-385,673-708), which also pins the decode semantics of src/cbor.rs:62-91.
//!
//! NOT COMPILED IN THIS REPOSITORY: the build image has no cargo/rustc (rust/README.md).
use anonymous_credit_tokens::{
    CreditToken, Error, IssuanceRequest, IssuanceResponse, Params, PreIssuance, PrivateKey, Refund, SpendProof,
};
use curve25519_dalek::Scalar;
use rand_chacha::ChaCha20Rng;
use rand_core::{CryptoRng, RngCore, SeedableRng};

/// Passes every request through to ChaCha20 and keeps a copy of the bytes handed out since the last `take()`.
struct Recorder {
    inner: ChaCha20Rng,
    log: Vec<u8>,
}
impl Recorder {
    fn take(&mut self) -> String {
        let s = hex(&self.log);
        self.log.clear();
        s
    }
}
impl RngCore for Recorder {
    fn next_u32(&mut self) -> u32 {
        let mut b = [0u8; 4];
        self.fill_bytes(&mut b);
        u32::from_le_bytes(b)
    }
    fn next_u64(&mut self) -> u64 {
        let mut b = [0u8; 8];
        self.fill_bytes(&mut b);
        u64::from_le_bytes(b)
    }
    fn fill_bytes(&mut self, dest: &mut [u8]) {
        self.inner.fill_bytes(dest);
        self.log.extend_from_slice(dest);
    }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand_core::Error> {
        self.fill_bytes(dest);
        Ok(())
    }
}
impl CryptoRng for Recorder {}

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}
fn err_code(e: &Error) -> u8 {
    // 1 + discriminant in declaration order (src/lib.rs:102-112) = the status byte of include/act_engine.h
    match e {
        Error::InvalidIssuanceRequestProof => 1,
        Error::InvalidIssuanceResponseProof => 2,
        Error::DoubleSpendError => 3,
        Error::InvalidRefundProof => 4,
        Error::InvalidRefundResponseProof => 5,
        Error::IdentityPointError => 6,
        Error::InvalidClientSpendProof => 7,
        Error::AmountTooBigError => 8,
        Error::ScalarOutOfRangeError => 9,
    }
}

/// Offset of the 32-byte payload of map key `key` in the canonical encodings ciborium produces for the flat records
/// (`a4`/`a5` map header, then `key 58 20 <32 bytes>` per field: src/cbor.rs:96-103,153-161,413-420,585-603).
fn flat_field(key: usize) -> usize {
    1 + (key - 1) * 35 + 3
}
/// Offset of the payload of a top-level 32-byte field of the SpendProof encoding (src/cbor.rs:216-268):
/// keys 1-4 precede the first array; 6-13 sit between com (key 5) and gamma0 (key 14); 16, 17 follow z (key 15).
fn proof_field(key: usize) -> usize {
    let arr = 1 + 2 + 128 * 34; // key, 98 80, 128 x (58 20 + 32)
    let zarr = 1 + 2 + 128 * (1 + 68); // key, 98 80, 128 x (82, two bstr)
    match key {
        1..=4 => 1 + (key - 1) * 35 + 3,
        6..=13 => 1 + 4 * 35 + arr + (key - 6) * 35 + 3,
        16 | 17 => 1 + 4 * 35 + arr + 8 * 35 + arr + zarr + (key - 16) * 35 + 3,
        _ => panic!("array field"),
    }
}
fn bump_scalar(bytes: &mut [u8], off: usize) {
    let mut s = [0u8; 32];
    s.copy_from_slice(&bytes[off..off + 32]);
    let v = Scalar::from_bytes_mod_order(s) + Scalar::ONE;
    bytes[off..off + 32].copy_from_slice(v.as_bytes());
}

struct Out {
    first: bool,
}
impl Out {
    fn item(&mut self, body: String) {
        if !self.first {
            println!(",");
        }
        self.first = false;
        print!("    {{{}}}", body);
    }
}

fn main() {
    let mut rng = Recorder { inner: ChaCha20Rng::from_seed([0x42u8; 32]), log: Vec::new() };
    let domain = ["test-org", "test-service", "test-env", "2024-01-01"]; // src/tests.rs:59
    let params = Params::new(domain[0], domain[1], domain[2], domain[3]);
    let key = PrivateKey::random(&mut rng);
    let key_rng = rng.take();
    println!("{{");
    println!("  \"source\": \"anonymous-credit-tokens 0.2.1 (real crate), ChaCha20 seed 0x42 x 32, rust/golden-dump\",");
    println!("  \"params\": [\"{}\", \"{}\", \"{}\", \"{}\"],", domain[0], domain[1], domain[2], domain[3]);
    println!("  \"private_key_cbor\": \"{}\",", hex(&key.to_cbor().unwrap()));
    println!("  \"public_key_cbor\": \"{}\",", hex(&key.public().to_cbor().unwrap()));
    println!("  \"private_key_rng\": \"{}\",", key_rng);
    println!("  \"calls\": [");
    let mut out = Out { first: true };

    for trip in 0..16u64 {
        let credits = 20 + 61 * trip; // 20 .. 935
        let charge = 1 + (7 * trip) % credits;
        let pre = PreIssuance::random(&mut rng);
        let pre_rng = rng.take();
        let request = pre.request(&params, &mut rng);
        let req_rng = rng.take();
        let req_cbor = request.to_cbor().unwrap();
        out.item(format!(
            "\"op\": \"request\", \"trip\": {}, \"preissuance_cbor\": \"{}\", \"preissuance_rng\": \"{}\", \"rng\": \"{}\", \"out_cbor\": \"{}\"",
            trip, hex(&pre.to_cbor().unwrap()), pre_rng, req_rng, hex(&req_cbor)
        ));
        // ---- issue: the honest request, then mutations of its CBOR (src/tests.rs:571-601,1934-1958)
        let c = Scalar::from(credits);
        let response = private_issue(&key, &params, &req_cbor, c, &mut rng, &mut out, trip, "valid").expect("valid request");
        {
            let mut m = req_cbor.clone();
            bump_scalar(&mut m, flat_field(3)); // k_bar + 1
            private_issue(&key, &params, &m, c, &mut rng, &mut out, trip, "k_bar+1");
            let mut m = req_cbor.clone();
            bump_scalar(&mut m, flat_field(2)); // gamma + 1
            private_issue(&key, &params, &m, c, &mut rng, &mut out, trip, "gamma+1");
            let mut m = req_cbor.clone();
            m[flat_field(4) + 31] |= 0xf0; // r_bar >= 2^252: a non-canonical encoding, reduced on decode (src/cbor.rs:80-91) to another scalar
            private_issue(&key, &params, &m, c, &mut rng, &mut out, trip, "r_bar high bits set");
        }
        let token = pre.to_credit_token(&params, key.public(), &request, &response).expect("issuance verifies");
        let token_cbor = token.to_cbor().unwrap();
        // ---- client check of the response and of tampered responses (src/tests.rs:692-720,825-848)
        let resp_cbor = response.to_cbor().unwrap();
        for (label, field) in [("valid", 0usize), ("e+1", 2), ("gamma+1", 3), ("z+1", 4), ("c+1", 5)] {
            let mut m = resp_cbor.clone();
            if field != 0 {
                bump_scalar(&mut m, flat_field(field));
            }
            let r = IssuanceResponse::from_cbor(&m).unwrap();
            let res = pre.to_credit_token(&params, key.public(), &request, &r);
            out.item(format!(
                "\"op\": \"issuance_check\", \"trip\": {}, \"label\": \"{}\", \"request_cbor\": \"{}\", \"response_cbor\": \"{}\", \"status\": {}",
                trip, label, hex(&req_cbor), hex(&m), res.as_ref().err().map(err_code).unwrap_or(0)
            ));
        }
        // ---- prove_spend (client) and refund (issuer): honest, then mutations of the proof's CBOR
        let (proof, prerefund) = token.prove_spend(&params, Scalar::from(charge), &mut rng);
        let prove_rng = rng.take();
        let proof_cbor = proof.to_cbor().unwrap();
        out.item(format!(
            "\"op\": \"prove_spend\", \"trip\": {}, \"token_cbor\": \"{}\", \"charge\": {}, \"rng\": \"{}\", \"out_cbor\": \"{}\", \"prerefund_cbor\": \"{}\"",
            trip, hex(&token_cbor), charge, prove_rng, hex(&proof_cbor), hex(&prerefund.to_cbor().unwrap())
        ));
        let refund = private_refund(&key, &params, &proof_cbor, &mut rng, &mut out, trip, "valid").expect("valid proof");
        for (label, key_no) in [("s+1", 2usize), ("k+1", 1), ("gamma+1", 6), ("e_bar+1", 7), ("r2_bar+1", 8), ("r3_bar+1", 9), ("c_bar+1", 10),
                                ("r_bar+1", 11), ("w00+1", 12), ("w01+1", 13), ("k_bar+1", 16), ("s_bar+1", 17)] {
            if (trip as usize + key_no) % 4 != 0 {
                continue; // a quarter of the classes per trip keeps the file small; 16 trips cover every class several times
            }
            let mut m = proof_cbor.clone();
            bump_scalar(&mut m, proof_field(key_no));
            private_refund(&key, &params, &m, &mut rng, &mut out, trip, label);
        }
        {
            let mut m = proof_cbor.clone();
            for b in &mut m[proof_field(3)..proof_field(3) + 32] {
                *b = 0; // A' = identity (src/tests.rs:851-873)
            }
            private_refund(&key, &params, &m, &mut rng, &mut out, trip, "A' identity");
            let mut m = proof_cbor.clone();
            let (a, b) = (proof_field(3), proof_field(4));
            for i in 0..32 {
                m.swap(a + i, b + i); // A' and B-bar swapped: both valid points
            }
            private_refund(&key, &params, &m, &mut rng, &mut out, trip, "A'/B swapped");
        }
        // ---- token tampering (src/tests.rs:1898-1927): a and e of the CreditToken replaced, then an honest prove_spend
        {
            let mut t = token_cbor.clone();
            let other = (&params_point(&mut rng)).to_vec();
            t[flat_field(1)..flat_field(1) + 32].copy_from_slice(&other);
            bump_scalar(&mut t, flat_field(2));
            let _ = rng.take();
            let bad = CreditToken::from_cbor(&t).unwrap();
            let (p2, _) = bad.prove_spend(&params, Scalar::ONE, &mut rng);
            let prng = rng.take();
            let p2c = p2.to_cbor().unwrap();
            out.item(format!(
                "\"op\": \"prove_spend\", \"trip\": {}, \"label\": \"tampered token\", \"token_cbor\": \"{}\", \"charge\": 1, \"rng\": \"{}\", \"out_cbor\": \"{}\"",
                trip, hex(&t), prng, hex(&p2c)
            ));
            private_refund(&key, &params, &p2c, &mut rng, &mut out, trip, "tampered token");
        }
        // ---- client check of the refund and of tampered refunds (src/tests.rs:781-822,1149-1231)
        let refund_cbor = refund.to_cbor().unwrap();
        for (label, field) in [("valid", 0usize), ("e+1", 2), ("gamma+1", 3), ("z+1", 4)] {
            let mut m = refund_cbor.clone();
            if field != 0 {
                bump_scalar(&mut m, flat_field(field));
            }
            let r = Refund::from_cbor(&m).unwrap();
            let res = prerefund.to_credit_token(&params, &proof, &r, key.public());
            out.item(format!(
                "\"op\": \"refund_check\", \"trip\": {}, \"label\": \"{}\", \"proof_cbor\": \"{}\", \"refund_cbor\": \"{}\", \"status\": {}",
                trip, label, hex(&proof_cbor), hex(&m), res.as_ref().err().map(err_code).unwrap_or(0)
            ));
        }
    }
    println!();
    println!("  ]");
    println!("}}");
}

/// A valid point encoding that is unrelated to the token: G * (random scalar), as compressed bytes.
fn params_point(rng: &mut Recorder) -> [u8; 32] {
    use curve25519_dalek::ristretto::RistrettoPoint;
    RistrettoPoint::mul_base(&Scalar::random(rng)).compress().to_bytes()
}

fn private_issue(key: &PrivateKey, params: &Params, req_cbor: &[u8], c: Scalar, rng: &mut Recorder, out: &mut Out, trip: u64, label: &str)
    -> Option<IssuanceResponse> {
    let request = IssuanceRequest::from_cbor(req_cbor).expect("structurally valid request");
    let res = key.issue(params, &request, c, &mut *rng);
    let drawn = rng.take();
    out.item(format!(
        "\"op\": \"issue\", \"trip\": {}, \"label\": \"{}\", \"request_cbor\": \"{}\", \"c\": \"{}\", \"rng\": \"{}\", \"status\": {}, \"out_cbor\": \"{}\"",
        trip, label, hex(req_cbor), hex(c.as_bytes()), drawn, res.as_ref().err().map(err_code).unwrap_or(0),
        res.as_ref().ok().map(|r| hex(&r.to_cbor().unwrap())).unwrap_or_default()
    ));
    res.ok()
}

fn private_refund(key: &PrivateKey, params: &Params, proof_cbor: &[u8], rng: &mut Recorder, out: &mut Out, trip: u64, label: &str) -> Option<Refund> {
    let proof = SpendProof::from_cbor(proof_cbor).expect("structurally valid proof");
    let res = key.refund(params, &proof, &mut *rng);
    let drawn = rng.take();
    out.item(format!(
        "\"op\": \"refund\", \"trip\": {}, \"label\": \"{}\", \"proof_cbor\": \"{}\", \"rng\": \"{}\", \"status\": {}, \"nullifier\": \"{}\", \"out_cbor\": \"{}\"",
        trip, label, hex(proof_cbor), drawn, res.as_ref().err().map(err_code).unwrap_or(0), hex(proof.nullifier().as_bytes()),
        res.as_ref().ok().map(|r| hex(&r.to_cbor().unwrap())).unwrap_or_default()
    ));
    res.ok()
}
