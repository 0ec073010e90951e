// Dumps one full PVSS round per group, produced by the reference itself (random coefficients and
// witnesses inside distribute_secret, participant.rs:160-286 / 1094-1274 / 1573-1717), as one JSON
// object on stdout.  All group elements / scalars are hex of Group::element_to_bytes /
// Group::scalar_to_bytes, i.e. exactly the bytes the reference hashes and keys its maps with.
//
// What the consumer can pin with it (tests/test_ref_vectors.py):
//   * verify_distribution_shares: recompute X_i, a1, a2, the framed transcript and the challenge from the
//     box alone -> must equal the dumped challenge (pins Horner-vs-reference X_i, framing, hash_to_scalar);
//   * extract_secret_share is deterministic given (sk, w): share, challenge, response must match;
//   * reconstruct: G^s path and the U mask -> the dumped secret.
use mpvss_rs::group::Group;
use mpvss_rs::groups::{ModpGroup, Ristretto255Group, Secp256k1Group};
use mpvss_rs::{string_to_secret, Participant};

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}

macro_rules! round {
    ($name:expr, $group_ty:ty, $n:expr, $t:expr) => {{
        let group = <$group_ty>::new();
        let secret = string_to_secret("Hello MPVSS Example.");
        let mut dealer = Participant::with_arc(group.clone());
        dealer.initialize();
        let mut ps = Vec::new();
        for _ in 0..$n {
            let mut p = Participant::with_arc(group.clone());
            p.initialize();
            ps.push(p);
        }
        let pks: Vec<_> = ps.iter().map(|p| p.publickey.clone()).collect();
        let dbox = dealer.distribute_secret(&secret, &pks, $t);
        assert!(ps[0].verify_distribution_shares(&dbox));
        let e = |x: &<$group_ty as Group>::Element| hex(&group.element_to_bytes(x));
        let s = |x: &<$group_ty as Group>::Scalar| hex(&group.scalar_to_bytes(x));
        let mut out = String::new();
        out.push_str(&format!("\"{}\": {{\"n\": {}, \"t\": {}, ", $name, $n, $t));
        out.push_str(&format!("\"secret\": \"{}\", ", hex(&secret.to_bytes_be().1)));
        out.push_str(&format!("\"U\": \"{}\", ", hex(&dbox.U.to_bytes_be().1)));
        out.push_str(&format!("\"challenge\": \"{}\", ", s(&dbox.challenge)));
        let list = |v: Vec<String>| format!("[{}]", v.iter().map(|x| format!("\"{}\"", x)).collect::<Vec<_>>().join(", "));
        out.push_str(&format!("\"commitments\": {}, ", list(dbox.commitments.iter().map(|c| e(c)).collect())));
        out.push_str(&format!("\"publickeys\": {}, ", list(pks.iter().map(|c| e(c)).collect())));
        let key = |pk: &<$group_ty as Group>::Element| group.element_to_bytes(pk);
        out.push_str(&format!("\"positions\": [{}], ",
            pks.iter().map(|pk| dbox.positions[&key(pk)].to_string()).collect::<Vec<_>>().join(", ")));
        out.push_str(&format!("\"shares\": {}, ", list(pks.iter().map(|pk| e(&dbox.shares[&key(pk)])).collect())));
        out.push_str(&format!("\"responses\": {}, ", list(pks.iter().map(|pk| s(&dbox.responses[&key(pk)])).collect())));
        out.push_str(&format!("\"private_keys\": {}, ", list(ps.iter().map(|p| s(&p.privatekey)).collect())));
        // decrypted shares with their proofs; the witness w is an argument of the reference API
        let mut ws = Vec::new();
        let mut boxes = Vec::new();
        for p in ps.iter() {
            let w = group.generate_private_key();
            let sb = p.extract_secret_share(&dbox, &p.privatekey, &w).expect("extract_secret_share");
            assert!(dealer.verify_share(&sb, &dbox, &p.publickey));
            ws.push(s(&w));
            boxes.push(sb);
        }
        out.push_str(&format!("\"extract_w\": {}, ", list(ws)));
        out.push_str(&format!("\"sharebox_share\": {}, ", list(boxes.iter().map(|b| e(&b.share)).collect())));
        out.push_str(&format!("\"sharebox_challenge\": {}, ", list(boxes.iter().map(|b| s(&b.challenge)).collect())));
        out.push_str(&format!("\"sharebox_response\": {}, ", list(boxes.iter().map(|b| s(&b.response)).collect())));
        let rec = dealer.reconstruct(&boxes[..$t as usize], &dbox).expect("reconstruct");
        assert_eq!(rec, secret);
        out.push_str(&format!("\"reconstructed\": \"{}\"}}", hex(&rec.to_bytes_be().1)));
        out
    }};
}

fn main() {
    let parts = vec![
        round!("modp", ModpGroup, 5, 3),
        round!("secp256k1", Secp256k1Group, 5, 3),
        round!("ristretto255", Ristretto255Group, 5, 3),
    ];
    println!("{{{}}}", parts.join(",\n "));
}
