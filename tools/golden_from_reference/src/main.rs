// Prints tests/golden/imgfprint_0.4.1.json: the 536-byte MultiHashFingerprint (hex) that imgfprint 0.4.1 -- the
// crate the reference calls at src/modality/image.rs:68-70 -- produces for the reference's own synthetic PNGs
// (src/server/tests.rs:227-235, benches/end_to_end.rs:77-85).  Needs cargo + network, so it cannot run in the
// offline development container; until its output is committed, tests/test_imgfprint_parity.py skips and image
// hash parity with imgfprint stays UNPINNED.
//   cargo run --release > ../../tests/golden/imgfprint_0.4.1.json
use imgfprint::{ImageFingerprinter, PreprocessConfig};

fn synthetic_png(w: u32, h: u32) -> Vec<u8> {
    let img = image::ImageBuffer::from_fn(w, h, |x, y| image::Rgb([(x % 256) as u8, (y % 256) as u8, 128u8]));
    let mut buf = Vec::new();
    img.write_to(&mut std::io::Cursor::new(&mut buf), image::ImageFormat::Png).unwrap();
    buf
}

fn main() {
    let mut items = Vec::new();
    for (w, h) in [(64u32, 64u32), (256, 256), (300, 200), (1024, 1024)] {
        let fp = ImageFingerprinter::fingerprint_with_preprocess(&synthetic_png(w, h), &PreprocessConfig::default()).unwrap();
        let hex: String = bytemuck::bytes_of(&fp).iter().map(|b| format!("{b:02x}")).collect();
        items.push(format!("  {{\"w\": {w}, \"h\": {h}, \"hex\": \"{hex}\"}}"));
    }
    println!("{{\"crate\": \"imgfprint 0.4.1\", \"images\": [\n{}\n]}}", items.join(",\n"));
}
