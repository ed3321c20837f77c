// NOT COMPILED HERE (no Go toolchain) — reviewed source; see host/go/README.md.
//
// Exposing a GPU-resident model to an agent container.  Follows the attachable injection (attachable.go:105-134, 161-182): the runner
// stages host files under the cell's metadata directory *before* the call, passes a value closure, and BuildContainerSpec turns it into
// oci.SpecOpts.  Needs one field on buildOpts (spec.go:178):  gpuWeights []GPUWeightsInjection  and, in BuildContainerSpec (spec.go:218),
//     for _, inj := range o.gpuWeights { specOpts = append(specOpts, withGPUWeights(inj)...) }

package ctr

import (
	"context"
	"fmt"
	"os"
	"path"
	"strings"

	"github.com/containerd/containerd/v2/core/containers"
	"github.com/containerd/containerd/v2/pkg/oci"
	runtimespec "github.com/opencontainers/runtime-spec/specs-go"
	"golang.org/x/sys/unix"
)

const (
	// GPUPoolContainerDir is where the staged manifest + IPC handle appear inside the container.
	GPUPoolContainerDir = "/run/kukeon/gpupool"
	envGPUPoolManifest  = "KUKEON_GPUPOOL_MANIFEST"
	envGPUPoolIPCHandle = "KUKEON_GPUPOOL_IPC_HANDLE"
	envGPUPoolDevice    = "KUKEON_GPUPOOL_DEVICE"
)

// GPUWeightsInjection is one mounted model.  HostDir holds manifest.json and ipc.handle, written by modelhub.Mount with the atomic
// tmp+fsync+rename helper (internal/metadata/metadata.go:105-140) the way secrets are staged (secrets.go:105-127).
type GPUWeightsInjection struct {
	Name    string // models[].name; empty = the single-model layout (no sub-directory, unsuffixed env names)
	HostDir string // <cell metadata dir>/<container>/gpupool[/<name>]
	Target  string // container path of the mount; empty = GPUPoolContainerDir
	Device  int    // CUDA ordinal whose pool the handle refers to
}

// WithGPUWeights appends one model to the container being built.  Repeatable: one option per models[] entry.
func WithGPUWeights(inj GPUWeightsInjection) BuildOption {
	return func(o *buildOpts) {
		o.gpuWeights = append(o.gpuWeights, inj)
	}
}

func gpuWeightsDest(inj GPUWeightsInjection) string {
	base := inj.Target
	if base == "" {
		base = GPUPoolContainerDir
	}
	if inj.Name == "" {
		return base
	}
	return path.Join(base, inj.Name)
}

// envSuffix: "llama-3.8b" -> "_LLAMA_3_8B" (POSIX environment names).
func envSuffix(name string) string {
	if name == "" {
		return ""
	}
	var b strings.Builder
	b.WriteByte('_')
	for _, r := range strings.ToUpper(name) {
		if (r >= 'A' && r <= 'Z') || (r >= '0' && r <= '9') {
			b.WriteRune(r)
		} else {
			b.WriteByte('_')
		}
	}
	return b.String()
}

// withGPUWeights renders the injection: a read-only bind of the staged directory (same shape bindVolumeMount emits, spec.go:526-543),
// the KUKEON_GPUPOOL_* environment (naming of kukeonDefaultEnv, spec.go:464-482) and the NVIDIA device nodes with their device-cgroup
// allow rules — the agent's CUDA runtime needs them to open the IPC handle, and the reference's spec builder emits no devices today.
func withGPUWeights(inj GPUWeightsInjection) []oci.SpecOpts {
	dest := gpuWeightsDest(inj)
	sfx := envSuffix(inj.Name)
	return []oci.SpecOpts{
		oci.WithMounts([]runtimespec.Mount{{
			Destination: dest,
			Source:      inj.HostDir,
			Type:        "bind",
			Options:     []string{"rbind", "ro"},
		}}),
		oci.WithEnv([]string{
			fmt.Sprintf("%s%s=%s/manifest.json", envGPUPoolManifest, sfx, dest),
			fmt.Sprintf("%s%s=%s/ipc.handle", envGPUPoolIPCHandle, sfx, dest),
			fmt.Sprintf("%s%s=%d", envGPUPoolDevice, sfx, inj.Device),
		}),
		withNvidiaDevices(inj.Device),
	}
}

// withNvidiaDevices adds /dev/nvidiactl, /dev/nvidia-uvm, /dev/nvidia-uvm-tools and /dev/nvidia<N> (those that exist) to Linux.Devices and
// allows them in the device cgroup.  Idempotent across several models on the same GPU: a node already present is skipped.
func withNvidiaDevices(device int) oci.SpecOpts {
	return func(_ context.Context, _ oci.Client, _ *containers.Container, s *runtimespec.Spec) error {
		if s.Linux == nil {
			s.Linux = &runtimespec.Linux{}
		}
		if s.Linux.Resources == nil {
			s.Linux.Resources = &runtimespec.LinuxResources{}
		}
		nodes := []string{"/dev/nvidiactl", "/dev/nvidia-uvm", "/dev/nvidia-uvm-tools", fmt.Sprintf("/dev/nvidia%d", device)}
		for _, p := range nodes {
			var st unix.Stat_t
			if err := unix.Stat(p, &st); err != nil {
				if os.IsNotExist(err) {
					continue
				}
				return fmt.Errorf("stat %s: %w", p, err)
			}
			if st.Mode&unix.S_IFMT != unix.S_IFCHR {
				continue
			}
			major, minor := int64(unix.Major(uint64(st.Rdev))), int64(unix.Minor(uint64(st.Rdev)))
			present := false
			for _, d := range s.Linux.Devices {
				if d.Path == p {
					present = true
					break
				}
			}
			if present {
				continue
			}
			mode := os.FileMode(0o666)
			var uid, gid uint32
			s.Linux.Devices = append(s.Linux.Devices, runtimespec.LinuxDevice{
				Path: p, Type: "c", Major: major, Minor: minor, FileMode: &mode, UID: &uid, GID: &gid,
			})
			maj, min := major, minor
			s.Linux.Resources.Devices = append(s.Linux.Resources.Devices, runtimespec.LinuxDeviceCgroup{
				Allow: true, Type: "c", Major: &maj, Minor: &min, Access: "rw",
			})
		}
		return nil
	}
}
