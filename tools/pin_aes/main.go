// pin_aes — produces the gnark-side evidence that pins this repository's AES-128/256 prove path (SURVEY.md §8c: the
// reference ships r1cs.aes* and vk.aes* but no pk.aes*, and its only acceptance test is "gnark's verifier accepts",
// libraries/core_test.go:174-260). NOT compiled or run in this repository (no Go toolchain on the build box); it is the one
// command a maintainer with Go runs once:
//
//	cp -r tools/pin_aes <checkout of reclaimprotocol/gnark-symmetric-crypto>/pin_aes
//	cd <checkout> && go run ./pin_aes -out /path/to/this/repo/tests/golden/external
//
// For bits in {128, 256} it
//  1. reads the reference's shipped circuits/generated/r1cs.aes<bits> (groth16.NewCS(ecc.BN254).ReadFrom, prove_impl.go:102-107),
//  2. runs groth16.Setup on it exactly as keygen.go:359-435 does and writes pk.aes<bits> / vk.aes<bits> (WriteTo),
//  3. proves the reference's own benchmark request (libraries/core_test.go:265 / :275) through the reference's
//     prover.InitAlgorithm + prover.Prove and checks it with groth16.Verify; writes request_<bits>.json / response_<bits>.json,
//  4. dumps what the BSB22 commitment challenge was derived from (gnark backend/groth16/bn254/prove.go:84-108): the bytes
//     handed to fr.Hash, the domain separation tag, and the resulting challenge -> challenge_<bits>.json.
//
// tests/test_external_pin.py activates as soon as these files exist: (i) this repo's parser loads the gnark pk, (ii) a GPU proof
// under the gnark pk is accepted under the gnark vk and the gnark proof is accepted by the GPU verifier, (iii) the challenge
// bytes agree.
package main

import (
	"bytes"
	"encoding/hex"
	"encoding/json"
	"flag"
	"fmt"
	"os"
	"path/filepath"

	"github.com/consensys/gnark-crypto/ecc"
	"github.com/consensys/gnark-crypto/ecc/bn254/fr"
	"github.com/consensys/gnark/backend/groth16"
	groth16_bn254 "github.com/consensys/gnark/backend/groth16/bn254"
	"github.com/consensys/gnark/backend/witness"
	"github.com/consensys/gnark/constraint"
	"github.com/consensys/gnark/frontend"

	aes_v2 "gnark-symmetric-crypto/circuits/aesV2"
	prover "gnark-symmetric-crypto/libraries/prover/impl"
)

type request struct {
	Cipher  string  `json:"cipher"`
	Key     []uint8 `json:"key"`
	Nonce   []uint8 `json:"nonce"`
	Counter uint32  `json:"counter"`
	Input   []uint8 `json:"input"`
}

type response struct {
	Proof struct {
		ProofJson []uint8 `json:"proofJson"`
	} `json:"proof"`
	PublicSignals []uint8 `json:"publicSignals"`
}

func must(err error) {
	if err != nil {
		panic(err)
	}
}

func fill(n int, v byte) []byte {
	b := make([]byte, n)
	for i := range b {
		b[i] = v
	}
	return b
}

// the reference's benchmark inputs: libraries/core_test.go:265 (AES-128) and :275 (AES-256)
var input128 = append([]byte{183, 4, 206, 60, 254, 21, 117, 9, 150, 227, 246, 245, 71, 101, 56, 67, 79, 93, 44, 163, 22, 89, 128, 55, 214,
	254, 228, 214, 89, 253, 176, 112, 138, 115, 93, 140, 194, 222, 104, 252, 49, 144, 91, 252}, make([]byte, 20)...)
var input256 = []byte{189, 250, 225, 242, 6, 46, 173, 203, 7, 166, 62, 139, 67, 150, 1, 155, 64, 122, 211, 198, 184, 203, 124, 194,
	99, 34, 127, 29, 236, 17, 232, 214, 154, 146, 78, 217, 254, 224, 208, 196, 55, 200, 23, 93, 90, 175, 240, 31,
	31, 225, 26, 15, 219, 156, 123, 21, 103, 98, 205, 87, 197, 22, 245, 158}

func publicWitness(req request, ct []byte) witness.Witness {
	w := aes_v2.AESWrapper{Key: make([]frontend.Variable, len(req.Key))}
	for i := range req.Key {
		w.Key[i] = req.Key[i]
	}
	for i := 0; i < 12; i++ {
		w.Nonce[i] = req.Nonce[i]
	}
	w.Counter = req.Counter
	for i := 0; i < 64; i++ {
		w.Plaintext[i] = req.Input[i]
		w.Ciphertext[i] = ct[i]
	}
	full, err := frontend.NewWitness(&w, ecc.BN254.ScalarField())
	must(err)
	pub, err := full.Public()
	must(err)
	return pub
}

func run(bits int, alg uint8, cipherName string, counter uint32, input []byte, out string) {
	r1csBytes, err := os.ReadFile(fmt.Sprintf("circuits/generated/r1cs.aes%d", bits))
	must(err)
	cs := groth16.NewCS(ecc.BN254)
	_, err = cs.ReadFrom(bytes.NewReader(r1csBytes))
	must(err)

	pk, vk, err := groth16.Setup(cs) // keygen.go:369,408
	must(err)
	var pkBuf, vkBuf bytes.Buffer
	_, err = pk.WriteTo(&pkBuf)
	must(err)
	_, err = vk.WriteTo(&vkBuf)
	must(err)
	must(os.WriteFile(filepath.Join(out, fmt.Sprintf("pk.aes%d", bits)), pkBuf.Bytes(), 0o644))
	must(os.WriteFile(filepath.Join(out, fmt.Sprintf("vk.aes%d", bits)), vkBuf.Bytes(), 0o644))

	if !prover.InitAlgorithm(alg, pkBuf.Bytes(), r1csBytes) {
		panic("InitAlgorithm failed")
	}
	req := request{Cipher: cipherName, Key: fill(bits/8, 2), Nonce: fill(12, 3), Counter: counter, Input: input}
	reqJSON, err := json.Marshal(req)
	must(err)
	resJSON := prover.Prove(reqJSON)
	var res response
	must(json.Unmarshal(resJSON, &res))
	must(os.WriteFile(filepath.Join(out, fmt.Sprintf("request_%d.json", bits)), reqJSON, 0o644))
	must(os.WriteFile(filepath.Join(out, fmt.Sprintf("response_%d.json", bits)), resJSON, 0o644))

	proof := groth16.NewProof(ecc.BN254)
	_, err = proof.ReadFrom(bytes.NewReader(res.Proof.ProofJson))
	must(err)
	must(groth16.Verify(proof, vk, publicWitness(req, res.PublicSignals)))

	// the BSB22 challenge as prove.go:84-108 derives it: fr.Hash(commitment.Marshal() || committed public values, dst, 1).
	// These circuits commit to private wires only (CommitmentInfo.PublicAndCommitmentCommitted is empty), so the hash input is
	// the uncompressed commitment alone.
	p := proof.(*groth16_bn254.Proof)
	if len(p.Commitments) != 1 {
		panic(fmt.Sprintf("expected one commitment, got %d", len(p.Commitments)))
	}
	hashInput := p.Commitments[0].Marshal()
	ch, err := fr.Hash(hashInput, []byte(constraint.CommitmentDst), 1)
	must(err)
	chBytes := ch[0].Bytes()
	dump := map[string]string{
		"hash_input_hex": hex.EncodeToString(hashInput),
		"dst":            constraint.CommitmentDst,
		"challenge_hex":  hex.EncodeToString(chBytes[:]),
		"gnark":          "github.com/consensys/gnark v0.11.0, gnark-crypto v0.14.0 (go.mod:8-9)",
	}
	dj, err := json.MarshalIndent(dump, "", " ")
	must(err)
	must(os.WriteFile(filepath.Join(out, fmt.Sprintf("challenge_%d.json", bits)), dj, 0o644))
	fmt.Printf("aes-%d: pk %d bytes, vk %d bytes, proof %d bytes, gnark verifier accepted\n", bits, pkBuf.Len(), vkBuf.Len(), len(res.Proof.ProofJson))
}

func main() {
	out := flag.String("out", "external", "directory for pk/vk/request/response/challenge files (this repo's tests/golden/external)")
	flag.Parse()
	must(os.MkdirAll(*out, 0o755))
	run(128, prover.AES_128, "aes-128-ctr", 2, input128, *out)
	run(256, prover.AES_256, "aes-256-ctr", 10, input256, *out)
}
