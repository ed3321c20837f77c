// NOT COMPILED HERE (no Go toolchain) — reviewed source; see host/go/README.md.
//
// Exposing a GPU-resident model to an agent container.  Follows the attachable injection (attachable.go:105-134, 161-182): the runner
// stages host files under the cell's metadata directory *before* the call, passes a value closure, and BuildContainerSpec turns it into
// oci.SpecOpts.  Needs one field on buildOpts (spec.go:178):  gpuWeights []GPUWeightsInjection  and, in BuildContainerSpec (spec.go:218),
//     for _, inj := range o.gpuWeights { specOpts = append(specOpts, withGPUWeights(inj)...) }

package ctr

import (
	"strconv"
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
	envGPUPoolDeviceUUID = "KUKEON_GPUPOOL_DEVICE_UUID" // "GPU-xxxxxxxx-...": the same string cudaGetDeviceProperties().uuid / nvidia-smi -L give inside the container
	envGPUPoolPCIBusID   = "KUKEON_GPUPOOL_PCI_BUS_ID"
)

// GPUWeightsInjection is one mounted model.  HostDir holds manifest.json and ipc.handle, written by modelhub.Mount with the atomic
// tmp+fsync+rename helper (internal/metadata/metadata.go:105-140) the way secrets are staged (secrets.go:105-127).
type GPUWeightsInjection struct {
	Name    string // models[].name; empty = the single-model layout (no sub-directory, unsuffixed env names)
	HostDir string // <cell metadata dir>/<container>/gpupool[/<name>]
	Target  string // container path of the mount; empty = GPUPoolContainerDir
	// Identity of the GPU whose pool the handle refers to, from gpupool.DeviceIdentity (kk_device_identity).  NOT the daemon's CUDA ordinal: ordinals
	// follow CUDA_DEVICE_ORDER / CUDA_VISIBLE_DEVICES and a container that sees only one /dev/nvidia<N> enumerates that GPU as ordinal 0.
	DeviceUUID string // "GPU-..."; the agent selects the CUDA device whose UUID matches
	PCIBusID   string // "dddd:bb:dd.f" (lower case): key of /proc/driver/nvidia/gpus/<id>/information, where the device-node minor is read
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
			fmt.Sprintf("%s%s=%s", envGPUPoolDeviceUUID, sfx, inj.DeviceUUID),
			fmt.Sprintf("%s%s=%s", envGPUPoolPCIBusID, sfx, inj.PCIBusID),
		}),
		withNvidiaDevices(inj.PCIBusID),
	}
}

// nvidiaDeviceMinor reads the minor number of the GPU's /dev/nvidia<N> node from the driver's table.  The CUDA ordinal is not that number
// (round-1 review): ordinals are sorted fastest-first and filtered by CUDA_VISIBLE_DEVICES, minors follow PCI enumeration.
func nvidiaDeviceMinor(pciBusID string) (int, error) {
	p := path.Join("/proc/driver/nvidia/gpus", strings.ToLower(pciBusID), "information")
	b, err := os.ReadFile(p)
	if err != nil {
		return 0, fmt.Errorf("read %s: %w", p, err)
	}
	for _, line := range strings.Split(string(b), "\n") {
		k, v, ok := strings.Cut(line, ":")
		if ok && strings.TrimSpace(k) == "Device Minor" {
			return strconv.Atoi(strings.TrimSpace(v))
		}
	}
	return 0, fmt.Errorf("%s: no Device Minor line", p)
}

// withNvidiaDevices adds /dev/nvidiactl, /dev/nvidia-uvm, /dev/nvidia-uvm-tools and the GPU's /dev/nvidia<minor> (those that exist) to Linux.Devices and
// allows them in the device cgroup.  Idempotent across several models on the same GPU: a node already present is skipped.
func withNvidiaDevices(pciBusID string) oci.SpecOpts {
	return func(_ context.Context, _ oci.Client, _ *containers.Container, s *runtimespec.Spec) error {
		device, err := nvidiaDeviceMinor(pciBusID)
		if err != nil {
			return err
		}
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
